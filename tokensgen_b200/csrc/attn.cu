// tcgen05 flash-attention forward, head_dim 64, non-causal.
//
//   O[q,:] = softmax(Q[q,:] K^T * scale) V          Q [BH, q_rows, 64], K/V [BH, kv_rows, 64] bf16
//
// One CTA = one (batch, head) and TWO 128-row query tiles, 384 threads:
//   warp 0       TMA producer : Q0,Q1 once; then a ring of {K block, V block} stages (128 kv rows each)
//   warp 1       MMA issuer   : S_t = Q_t K^T  (SS, 128x128x64)   and   O_t += P_t V  (TS: P from TMEM, V MN-major)
//   warp 2       TMEM allocator
//   warps 4..7   softmax for tile 0 (thread = one query row; no cross-thread reductions)
//   warps 8..11  softmax for tile 1
// The two tiles ping-pong: while one tile's rows are in softmax (MUFU-bound at hd=64), the tensor core runs the
// other tile's MMAs.  TMEM (512 columns): S_t 128 fp32 columns, O_t 64, P_t 64 (bf16 pairs) for t = 0,1.  P_t has its
// own columns so that S_t can be overwritten by the NEXT block's QK^T as soon as the softmax threads have pulled S_t
// into registers (s_free barrier) — the QK^T round trip is then off the softmax critical path.  The PV MMA reads P_t
// straight from TMEM.  O_t is rescaled lazily (only when the running max grows by > 2^8), which keeps the exact result
// because P and the row sum always share the same reference max.
#include <cuda_bf16.h>

#include <string>

#include "common.h"
#include "ptx.cuh"

namespace tg {

constexpr int AT_BLOCK_Q = 128;
constexpr int AT_BLOCK_KV = 128;
constexpr int AT_D = 64;
constexpr int AT_STAGES = 4;
constexpr int AT_TILE_BYTES = 128 * 64 * 2;  // 16 KB: Q tile, K block, V block
constexpr int AT_SMEM_BYTES = 2 * AT_TILE_BYTES + AT_STAGES * 2 * AT_TILE_BYTES + 256 + 1024;
constexpr int AT_THREADS = 384;

struct AttnParams {
    int q_rows, kv_rows;
    int H;
    __nv_bfloat16* out;
    int64_t out_rows_alloc, out_row0;
    float scale_log2;  // softmax_scale * log2(e)
    int accumulate;
    float out_scale;
    long long* trace;  // developer timeline (tools/attn_trace.py): [16 softmax warps][n_blocks][6] clock64 stamps of CTA (0,0)
    int n_pass;        // v3: 1, or 2 = a second (Q2, K2, V2) problem over the same query rows, accumulated into the same output
    int kv_rows2;      // v3 pass 1
    float out_scale2;  // v3 pass 1: out += out_scale2 * attn2
    int spec;          // v3: speculative softmax reference (block 0's row max, never updated) with an exact in-kernel redo
    int mutex;    // v2: the two tiles take turns on the MUFU pipe (named-barrier hand-off) instead of sharing it
    int stagger;  // v2: cycles by which tile 1 starts after tile 0 (keeps the two tiles' softmax phases interleaved)
    // sequence-parallel scatter of the output rows (tg_attn_fwd_sp): sp_world > 0 -> row g of a batch goes to its owner rank
    int sp_world, sp_chunk, sp_rows, sp_h_total, sp_head0;
    __nv_bfloat16* sp_peer[TG_MAX_PEERS];
};

// Address of output row (batch b, query q_row) of local head h: token-major [B, rows, H*64] locally, or — sequence
// parallel — the row's owner rank's [B, rows_local(owner), H_total*64] buffer over NVLink (second Ulysses all-to-all).
__device__ __forceinline__ __nv_bfloat16* attn_out_row(const AttnParams& p, int b, int h, int q_row) {
    if (p.sp_world > 0) {
        const int g = int(p.out_row0) + q_row;
        const int owner = min(g / p.sp_chunk, p.sp_world - 1);
        const int rl = owner == p.sp_world - 1 ? p.sp_rows - owner * p.sp_chunk : p.sp_chunk;
        return p.sp_peer[owner] + ((int64_t(b) * rl + (g - owner * p.sp_chunk)) * p.sp_h_total + p.sp_head0 + h) * AT_D;
    }
    return p.out + ((int64_t(b) * p.out_rows_alloc + p.out_row0 + q_row) * p.H + h) * AT_D;
}

#ifdef TG_DEVELOPER  // kernel generations 1 and 2: kept for tools/attn_time.py / attn_trace.py comparisons, not shipped
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t q_smem = smem_base;                      // 2 tiles
    const uint32_t kv_smem = smem_base + 2 * AT_TILE_BYTES;  // stages x {K, V}
    const uint32_t bar_base = kv_smem + AT_STAGES * 2 * AT_TILE_BYTES;
    const uint32_t q_full = bar_base;
    auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
    auto kv_empty = [&](int s) { return bar_base + 8u * (1 + AT_STAGES + s); };
    auto s_full = [&](int t) { return bar_base + 8u * (1 + 2 * AT_STAGES + t); };
    auto p_full = [&](int t) { return bar_base + 8u * (3 + 2 * AT_STAGES + t); };
    auto o_done = [&](int t) { return bar_base + 8u * (5 + 2 * AT_STAGES + t); };
    auto s_free = [&](int t) { return bar_base + 8u * (7 + 2 * AT_STAGES + t); };
    const uint32_t tmem_slot = bar_base + 8u * (9 + 2 * AT_STAGES);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int bh = blockIdx.y;
    const int q0 = blockIdx.x * (2 * AT_BLOCK_Q);
    const int n_blocks = (p.kv_rows + AT_BLOCK_KV - 1) / AT_BLOCK_KV;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_k);
        tma_prefetch_desc(&tmap_v);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < AT_STAGES; ++s) {
            mbar_init(kv_full(s), 1);
            mbar_init(kv_empty(s), 1);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(s_full(t), 1);
            mbar_init(p_full(t), 4);  // one arrive per softmax warp of the tile
            mbar_init(s_free(t), 4);
            mbar_init(o_done(t), 1);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    // TMEM columns: S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384) P0 [384,448) P1 [448,512)
    auto s_col = [&](int t) { return uint32_t(t * 128); };
    auto o_col = [&](int t) { return uint32_t(256 + t * 64); };
    auto p_col = [&](int t) { return uint32_t(384 + t * 64); };

    if (warp < 4) {
        reg_dealloc<80>();
        if (warp == 0) {
            // -------------------------------------------------------------- TMA producer
            if (lane == 0) {
                mbar_arrive_expect_tx(q_full, 2 * AT_TILE_BYTES);
                tma_load_3d(q_smem, &tmap_q, q_full, 0, q0, bh);
                tma_load_3d(q_smem + AT_TILE_BYTES, &tmap_q, q_full, 0, q0 + AT_BLOCK_Q, bh);
                int stage = 0;
                uint32_t phase = 0;
                for (int j = 0; j < n_blocks; ++j) {
                    mbar_wait(kv_empty(stage), phase ^ 1u, 0x201);
                    const uint32_t ks = kv_smem + stage * 2 * AT_TILE_BYTES;
                    mbar_arrive_expect_tx(kv_full(stage), 2 * AT_TILE_BYTES);
                    tma_load_3d(ks, &tmap_k, kv_full(stage), 0, j * AT_BLOCK_KV, bh);
                    tma_load_3d(ks + AT_TILE_BYTES, &tmap_v, kv_full(stage), 0, j * AT_BLOCK_KV, bh);
                    if (++stage == AT_STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        } else if (warp == 1) {
            // -------------------------------------------------------------- MMA issuer
            if (lane == 0) {
                constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, false, false);
                constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, false, true);  // V is MN-major
                auto issue_qk = [&](int t, int stage) {
                    const uint32_t qs = q_smem + t * AT_TILE_BYTES;
                    const uint32_t ks = kv_smem + stage * 2 * AT_TILE_BYTES;
#pragma unroll
                    for (int k = 0; k < AT_D / 16; ++k) {
                        const uint64_t da = make_smem_desc_sw128(qs + k * 32, 16, 1024);
                        const uint64_t db = make_smem_desc_sw128(ks + k * 32, 16, 1024);
                        umma_ss(tmem_base + s_col(t), da, db, idesc_qk, k != 0 ? 1u : 0u);
                    }
                    umma_commit(s_full(t));
                };
                auto issue_pv = [&](int t, int stage, bool first) {
                    const uint32_t vs = kv_smem + stage * 2 * AT_TILE_BYTES + AT_TILE_BYTES;
#pragma unroll
                    for (int k = 0; k < AT_BLOCK_KV / 16; ++k) {
                        // 16 kv rows per MMA = two 8-row swizzle groups (SBO 1024 B apart); N = 64 is one MN atom.
                        const uint64_t db = make_smem_desc_sw128(vs + k * 2048, 1024, 1024);
                        umma_ts(tmem_base + o_col(t), tmem_base + p_col(t) + uint32_t(k * 8), db, idesc_pv,
                                (first && k == 0) ? 0u : 1u);
                    }
                };
                mbar_wait(q_full, 0, 0x202);
                mbar_wait(kv_full(0), 0, 0x203);
                tc_fence_after();
                issue_qk(0, 0);
                issue_qk(1, 0);
                int stage = 0;
                uint32_t phase = 0;
                for (int j = 0; j < n_blocks; ++j) {
                    int nstage = stage + 1;
                    uint32_t nphase = phase;
                    if (nstage == AT_STAGES) {
                        nstage = 0;
                        nphase ^= 1u;
                    }
                    const bool has_next = (j + 1 < n_blocks);
                    for (int t = 0; t < 2; ++t) {
                        if (has_next) {
                            // S_t(j) has been pulled into registers: overwrite it with the next block's scores now
                            mbar_wait(s_free(t), uint32_t(j & 1), 0x209);
                            if (t == 0) mbar_wait(kv_full(nstage), nphase, 0x205);
                            tc_fence_after();
                            issue_qk(t, nstage);
                        }
                        mbar_wait(p_full(t), uint32_t(j & 1), 0x204);
                        tc_fence_after();
                        issue_pv(t, stage, j == 0);
                        if (t == 1) umma_commit(kv_empty(stage));
                        umma_commit(o_done(t));
                    }
                    stage = nstage;
                    phase = nphase;
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax warpgroups
        reg_alloc<208>();
        const int t = (warp - 4) >> 2;  // tile
        const int wq = warp & 3;        // TMEM lane quarter
        const int row_in_tile = wq * 32 + lane;
        const int q_row = q0 + t * AT_BLOCK_Q + row_in_tile;
        const uint32_t lane_base = tmem_base + (uint32_t(wq * 32) << 16);
        const uint32_t s_addr = lane_base + s_col(t);
        const uint32_t o_addr = lane_base + o_col(t);
        const uint32_t p_addr = lane_base + p_col(t);
        const float c = p.scale_log2;
        float m_ref = -INFINITY;  // reference max (raw score units)
        float l = 0.f;            // running sum of exp2((s - m_ref) * c)

        for (int j = 0; j < n_blocks; ++j) {
            mbar_wait(s_full(t), uint32_t(j & 1), 0x206);
            tc_fence_after();
            float s[128];
            {
                // issue all four TMEM loads, then one wait: the loads overlap each other
                uint32_t r0[32], r1[32], r2[32], r3[32];
                tmem_ld32(s_addr, r0);
                tmem_ld32(s_addr + 32, r1);
                tmem_ld32(s_addr + 64, r2);
                tmem_ld32(s_addr + 96, r3);
                tmem_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(s_free(t));
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    s[i] = __uint_as_float(r0[i]);
                    s[32 + i] = __uint_as_float(r1[i]);
                    s[64 + i] = __uint_as_float(r2[i]);
                    s[96 + i] = __uint_as_float(r3[i]);
                }
            }
            const int valid = p.kv_rows - j * AT_BLOCK_KV;
            if (valid < AT_BLOCK_KV) {
#pragma unroll
                for (int i = 0; i < 128; ++i)
                    if (i >= valid) s[i] = -INFINITY;
            }
            // row max with 8 independent chains (a single chain of 128 dependent FMNMX costs ~500 cycles of latency)
            float pm[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) pm[k] = fmaxf(s[2 * k], s[2 * k + 1]);
#pragma unroll
            for (int i = 16; i < 128; i += 16)
#pragma unroll
                for (int k = 0; k < 8; ++k) pm[k] = fmaxf(pm[k], fmaxf(s[i + 2 * k], s[i + 2 * k + 1]));
            const float mx = fmaxf(fmaxf(fmaxf(pm[0], pm[1]), fmaxf(pm[2], pm[3])), fmaxf(fmaxf(pm[4], pm[5]), fmaxf(pm[6], pm[7])));

            const bool need = (mx - m_ref) * c > 8.0f;  // true on the first block (m_ref = -inf)
            if (j == 0) {
                m_ref = mx;
            } else {
                // PV_t(j-1) must have landed before P_t is overwritten or O_t rescaled (almost always already true)
                mbar_wait(o_done(t), uint32_t((j - 1) & 1), 0x207);
                tc_fence_after();
                if (__any_sync(0xffffffffu, need)) {
                    float alpha = 1.0f;
                    if (need) {
                        alpha = fast_exp2((m_ref - mx) * c);
                        m_ref = mx;
                        l *= alpha;
                    }
#pragma unroll
                    for (int cc = 0; cc < 2; ++cc) {
                        uint32_t r[32];
                        tmem_ld32(o_addr + cc * 32, r);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                        tmem_st32(o_addr + cc * 32, r);
                    }
                }
            }
            const float mc = m_ref * c;
            float ps[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) ps[k] = 0.f;
#pragma unroll
            for (int i = 0; i < 128; i += 8)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    s[i + k] = fast_exp2(fmaf(s[i + k], c, -mc));
                    ps[k] += s[i + k];
                }
            l += ((ps[0] + ps[1]) + (ps[2] + ps[3])) + ((ps[4] + ps[5]) + (ps[6] + ps[7]));
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                uint32_t r[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i] = pack_bf16x2(s[cc * 64 + 2 * i], s[cc * 64 + 2 * i + 1]);
                tmem_st32(p_addr + cc * 32, r);
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full(t));
        }

        // ---- epilogue: O / l -> global
        mbar_wait(o_done(t), uint32_t((n_blocks - 1) & 1), 0x208);
        tc_fence_after();
        const float inv_l = 1.0f / l;
        const bool store = q_row < p.q_rows;
        const int b = bh / p.H, h = bh - b * p.H;
        __nv_bfloat16* o_ptr =
            attn_out_row(p, b, h, q_row);
        uint4 prev[8];
        if (store && p.accumulate) {  // all loads of the read-modify-write first (see gemm.cu: epi_gate_residual)
#pragma unroll
            for (int i = 0; i < 8; ++i) prev[i] = reinterpret_cast<const uint4*>(o_ptr)[i];
        }
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            uint32_t r[32];
            tmem_ld32(o_addr + cc * 32, r);
            tmem_wait_ld();
            if (store) {
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    float f[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) f[k] = __uint_as_float(r[i + k]) * inv_l;
                    if (p.accumulate) {
                        const uint4 old = prev[cc * 4 + i / 8];
                        f[0] = fmaf(p.out_scale, f[0], bf16_lo(old.x)); f[1] = fmaf(p.out_scale, f[1], bf16_hi(old.x));
                        f[2] = fmaf(p.out_scale, f[2], bf16_lo(old.y)); f[3] = fmaf(p.out_scale, f[3], bf16_hi(old.y));
                        f[4] = fmaf(p.out_scale, f[4], bf16_lo(old.z)); f[5] = fmaf(p.out_scale, f[5], bf16_hi(old.z));
                        f[6] = fmaf(p.out_scale, f[6], bf16_lo(old.w)); f[7] = fmaf(p.out_scale, f[7], bf16_hi(old.w));
                    }
                    uint4 v;
                    v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]);
                    v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
                    reinterpret_cast<uint4*>(o_ptr + cc * 32 + i)[0] = v;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}


// =====================================================================================================================
// v2: same tiling (one CTA = one (batch, head) x two 128-row query tiles, 128-row KV blocks, 4-stage TMA ring) with
//   * 16 softmax warps (4 per SM sub-partition instead of 2): every query row is shared by two threads, each owning 64
//     of the block's 128 score columns; they exchange their partial row max through shared memory (one 64-thread named
//     barrier per block).  The extra warps hide the TMEM-load / barrier / MUFU latencies that left the MUFU pipe at 62 %.
//   * one MMA-issuing thread PER TILE (warps 1 and 3) so that neither tile's QK^T / PV issue ever queues behind a wait
//     on the other tile's softmax.
//   * the wait for the previous PV (o_done) moved from before the exponentials to just before P is overwritten.
//   * EMU of every 8 exponentials evaluated on the FMA pipe (Cody-Waite split + degree-3 polynomial, max rel. error
//     8.8e-5, far below the bf16 rounding of P) instead of MUFU.EX2 (16 results/clk/SM is the binding limit at hd = 64).
constexpr int A2_THREADS = 640;
constexpr int A2_XCH_BYTES = 2 * 2 * 2 * 128 * 4;  // [parity][tile][half][row] fp32
constexpr int A2_SMEM_BYTES = 2 * AT_TILE_BYTES + AT_STAGES * 2 * AT_TILE_BYTES + A2_XCH_BYTES + 256 + 1024;

__device__ __forceinline__ float ex2_poly(float x) {
    x = fmaxf(x, -127.0f);
    float t;
    asm("add.rm.ftz.f32 %0, %1, 0f4B400000;" : "=f"(t) : "f"(x));  // x + 1.5*2^23, rounded down: low mantissa bits = floor(x)
    const float fl = t - 12582912.0f;
    const float f = x - fl;  // [0, 1)
    float p = fmaf(f, 0.077119089663028717f, 0.227564394474029541f);
    p = fmaf(p, f, 0.695146143436431885f);
    p = fmaf(p, f, 1.0f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

template <int EMU, bool MUTEX, bool TRACE, bool PACKED>
__global__ void __launch_bounds__(A2_THREADS, 1)
attn2_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t q_smem = smem_base;
    const uint32_t kv_smem = smem_base + 2 * AT_TILE_BYTES;
    const uint32_t xch_smem = kv_smem + AT_STAGES * 2 * AT_TILE_BYTES;
    const uint32_t bar_base = xch_smem + A2_XCH_BYTES;
    const uint32_t q_full = bar_base;
    auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
    auto kv_empty = [&](int s) { return bar_base + 8u * (1 + AT_STAGES + s); };
    auto s_full = [&](int t) { return bar_base + 8u * (1 + 2 * AT_STAGES + t); };
    auto p_full = [&](int t) { return bar_base + 8u * (3 + 2 * AT_STAGES + t); };
    auto o_done = [&](int t) { return bar_base + 8u * (5 + 2 * AT_STAGES + t); };
    auto s_free = [&](int t) { return bar_base + 8u * (7 + 2 * AT_STAGES + t); };
    const uint32_t tmem_slot = bar_base + 8u * (9 + 2 * AT_STAGES);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    float* xch = reinterpret_cast<float*>(smem_raw + (xch_smem - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int bh = blockIdx.y;
    const int q0 = blockIdx.x * (2 * AT_BLOCK_Q);
    const int n_blocks = (p.kv_rows + AT_BLOCK_KV - 1) / AT_BLOCK_KV;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_k);
        tma_prefetch_desc(&tmap_v);
    }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < AT_STAGES; ++s) {
            mbar_init(kv_full(s), 1);
            mbar_init(kv_empty(s), 2);  // one commit per tile's MMA thread
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(s_full(t), 1);
            mbar_init(p_full(t), 8);  // one arrive per softmax warp of the tile
            mbar_init(s_free(t), 8);
            mbar_init(o_done(t), 1);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    auto s_col = [&](int t) { return uint32_t(t * 128); };
    auto o_col = [&](int t) { return uint32_t(256 + t * 64); };
    auto p_col = [&](int t) { return uint32_t(384 + t * 64); };

    if (warp < 4) {
        reg_dealloc<40>();
        if (warp == 0) {
            // -------------------------------------------------------------- TMA producer
            if (lane == 0) {
                mbar_arrive_expect_tx(q_full, 2 * AT_TILE_BYTES);
                tma_load_3d(q_smem, &tmap_q, q_full, 0, q0, bh);
                tma_load_3d(q_smem + AT_TILE_BYTES, &tmap_q, q_full, 0, q0 + AT_BLOCK_Q, bh);
                int stage = 0;
                uint32_t phase = 0;
                for (int j = 0; j < n_blocks; ++j) {
                    mbar_wait(kv_empty(stage), phase ^ 1u, 0x301);
                    const uint32_t ks = kv_smem + stage * 2 * AT_TILE_BYTES;
                    mbar_arrive_expect_tx(kv_full(stage), 2 * AT_TILE_BYTES);
                    tma_load_3d(ks, &tmap_k, kv_full(stage), 0, j * AT_BLOCK_KV, bh);
                    tma_load_3d(ks + AT_TILE_BYTES, &tmap_v, kv_full(stage), 0, j * AT_BLOCK_KV, bh);
                    if (++stage == AT_STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        } else if (warp == 1 || warp == 3) {
            // -------------------------------------------------------------- MMA issuer of tile t
            if (lane == 0) {
                const int t = warp >> 1;
                constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, false, false);
                constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, false, true);  // V is MN-major
                const uint32_t qs = q_smem + t * AT_TILE_BYTES;
                auto issue_qk = [&](int stage) {
                    const uint32_t ks = kv_smem + stage * 2 * AT_TILE_BYTES;
#pragma unroll
                    for (int k = 0; k < AT_D / 16; ++k) {
                        const uint64_t da = make_smem_desc_sw128(qs + k * 32, 16, 1024);
                        const uint64_t db = make_smem_desc_sw128(ks + k * 32, 16, 1024);
                        umma_ss(tmem_base + s_col(t), da, db, idesc_qk, k != 0 ? 1u : 0u);
                    }
                    umma_commit(s_full(t));
                };
                auto issue_pv = [&](int stage, bool first) {
                    const uint32_t vs = kv_smem + stage * 2 * AT_TILE_BYTES + AT_TILE_BYTES;
#pragma unroll
                    for (int k = 0; k < AT_BLOCK_KV / 16; ++k) {
                        const uint64_t db = make_smem_desc_sw128(vs + k * 2048, 1024, 1024);
                        umma_ts(tmem_base + o_col(t), tmem_base + p_col(t) + uint32_t(k * 8), db, idesc_pv,
                                (first && k == 0) ? 0u : 1u);
                    }
                };
                mbar_wait(q_full, 0, 0x302);
                mbar_wait(kv_full(0), 0, 0x303);
                if (t == 1 && p.stagger > 0) {
                    // The two tiles only interact through the shared MUFU pipe, which preserves whatever phase offset
                    // they start with: started together, both sit in their latency-bound phase (TMEM
                    // load, row max, barriers) at the same time and the MUFU pipe idles for that long every block.
                    const long long t0 = clock64();
                    while (clock64() - t0 < p.stagger) {}
                }
                tc_fence_after();
                issue_qk(0);
                int stage = 0;
                uint32_t phase = 0;
                for (int j = 0; j < n_blocks; ++j) {
                    int nstage = stage + 1;
                    uint32_t nphase = phase;
                    if (nstage == AT_STAGES) {
                        nstage = 0;
                        nphase ^= 1u;
                    }
                    if (j + 1 < n_blocks) {
                        mbar_wait(s_free(t), uint32_t(j & 1), 0x309);  // S_t(j) is in registers: overwrite it
                        mbar_wait(kv_full(nstage), nphase, 0x305);
                        tc_fence_after();
                        issue_qk(nstage);
                    }
                    mbar_wait(p_full(t), uint32_t(j & 1), 0x304);
                    tc_fence_after();
                    issue_pv(stage, j == 0);
                    umma_commit(kv_empty(stage));
                    umma_commit(o_done(t));
                    stage = nstage;
                    phase = nphase;
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax warps
        reg_alloc<104>();
        const int sw = warp - 4;
        const int t = sw >> 3;          // tile
        const int half = (sw >> 2) & 1;  // which 64 of the block's 128 score columns
        const int wq = warp & 3;         // TMEM lane quarter
        const int row_in_tile = wq * 32 + lane;
        const int q_row = q0 + t * AT_BLOCK_Q + row_in_tile;
        const uint32_t lane_base = tmem_base + (uint32_t(wq * 32) << 16);
        const uint32_t s_addr = lane_base + s_col(t) + uint32_t(half * 64);
        const uint32_t o_addr = lane_base + o_col(t) + uint32_t(half * 32);
        const uint32_t p_addr = lane_base + p_col(t) + uint32_t(half * 32);
        const int pair_bar = 1 + t * 4 + wq;
        const float c = p.scale_log2;
        float m_ref = -INFINITY;
        float l = 0.f;

        for (int j = 0; j < n_blocks; ++j) {
            const bool tr = TRACE && p.trace != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0;
            long long* trp = p.trace + (int64_t(sw) * n_blocks + j) * 6;
            if (tr) trp[0] = clock64();
            mbar_wait(s_full(t), uint32_t(j & 1), 0x306);
            tc_fence_after();
            if (tr) trp[1] = clock64();
            float s[64];
            {
                uint32_t r0[32], r1[32];
                tmem_ld32(s_addr, r0);
                tmem_ld32(s_addr + 32, r1);
                tmem_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(s_free(t));
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    s[i] = __uint_as_float(r0[i]);
                    s[32 + i] = __uint_as_float(r1[i]);
                }
            }
            if (tr) trp[2] = clock64();
            const int valid = p.kv_rows - j * AT_BLOCK_KV - half * 64;
            if (valid < 64) {
#pragma unroll
                for (int i = 0; i < 64; ++i)
                    if (i >= valid) s[i] = -INFINITY;
            }
            float pm[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) pm[k] = fmaxf(s[2 * k], s[2 * k + 1]);
#pragma unroll
            for (int i = 8; i < 64; i += 8)
#pragma unroll
                for (int k = 0; k < 4; ++k) pm[k] = fmaxf(pm[k], fmaxf(s[i + 2 * k], s[i + 2 * k + 1]));
            float mx = fmaxf(fmaxf(pm[0], pm[1]), fmaxf(pm[2], pm[3]));
            {
                float* slot = xch + ((((j & 1) * 2 + t) * 2) * 128) + row_in_tile;
                slot[half * 128] = mx;
                named_bar_sync(pair_bar, 64);
                mx = fmaxf(mx, slot[(half ^ 1) * 128]);
            }
            if (tr) trp[3] = clock64();
            bool waited = false;
            if (j == 0) {
                m_ref = mx;
            } else {
                const bool need = (mx - m_ref) * c > 8.0f;
                if (__any_sync(0xffffffffu, need)) {
                    mbar_wait(o_done(t), uint32_t((j - 1) & 1), 0x307);
                    tc_fence_after();
                    waited = true;
                    float alpha = 1.0f;
                    if (need) {
                        alpha = fast_exp2((m_ref - mx) * c);
                        m_ref = mx;
                        l *= alpha;
                    }
                    uint32_t r[32];
                    tmem_ld32(o_addr, r);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                    tmem_st32(o_addr, r);
                }
            }
            const float mc = m_ref * c;
            float ps[4] = {0.f, 0.f, 0.f, 0.f};
            // MUFU hand-off: tile t may start its exponentials once tile 1-t has finished its own (ids 9, 10; 256 waiting
            // + 256 arriving threads).  Shared fairly, both tiles would sit in their latency-bound phases (S load, row max
            // exchange, P store) at the same time and leave the MUFU pipe idle for that long in every block.
            if (MUTEX && (j > 0 || t == 1)) asm volatile("bar.sync %0, 512;" ::"r"(10 - t) : "memory");
            if constexpr (PACKED) {
                const uint64_t c2 = pack_f32x2(c, c), nmc2 = pack_f32x2(-mc, -mc);
                uint64_t ps2[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
                for (int i = 0; i < 64; i += 8)
#pragma unroll
                    for (int k = 0; k < 8; k += 2) {
                        const uint64_t x2 = fma_f32x2(pack_f32x2(s[i + k], s[i + k + 1]), c2, nmc2);
                        s[i + k] = (k < EMU) ? ex2_poly(f32x2_lo(x2)) : fast_exp2(f32x2_lo(x2));
                        s[i + k + 1] = (k + 1 < EMU) ? ex2_poly(f32x2_hi(x2)) : fast_exp2(f32x2_hi(x2));
                        ps2[k >> 1] = add_f32x2(ps2[k >> 1], pack_f32x2(s[i + k], s[i + k + 1]));
                    }
                const uint64_t a = add_f32x2(add_f32x2(ps2[0], ps2[1]), add_f32x2(ps2[2], ps2[3]));
                l += f32x2_lo(a) + f32x2_hi(a);
            } else {
#pragma unroll
                for (int i = 0; i < 64; i += 8)
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const float x = fmaf(s[i + k], c, -mc);
                        s[i + k] = (k < EMU) ? ex2_poly(x) : fast_exp2(x);
                        ps[k & 3] += s[i + k];
                    }
                l += (ps[0] + ps[1]) + (ps[2] + ps[3]);
            }
            if (MUTEX) asm volatile("bar.arrive %0, 512;" ::"r"(9 + t) : "memory");
            uint32_t r[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = pack_bf16x2(s[2 * i], s[2 * i + 1]);
            if (tr) trp[4] = clock64();
            if (j > 0 && !waited) {  // PV_t(j-1) must have read P_t before it is overwritten (long done by now)
                mbar_wait(o_done(t), uint32_t((j - 1) & 1), 0x30a);
                tc_fence_after();
            }
            tmem_st32(p_addr, r);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full(t));
            if (tr) trp[5] = clock64();
        }

        if (MUTEX && t == 0) asm volatile("bar.sync 10, 512;" ::: "memory");  // consume tile 1's last hand-off
        // ---- epilogue: O / l -> global (this thread: 32 of the row's 64 output columns)
        {
            float* slot = xch + ((((n_blocks & 1) * 2 + t) * 2) * 128) + row_in_tile;
            slot[half * 128] = l;
            named_bar_sync(pair_bar, 64);
            l += slot[(half ^ 1) * 128];
        }
        mbar_wait(o_done(t), uint32_t((n_blocks - 1) & 1), 0x308);
        tc_fence_after();
        const float inv_l = 1.0f / l;
        const bool store = q_row < p.q_rows;
        const int b = bh / p.H, h = bh - b * p.H;
        __nv_bfloat16* o_ptr =
            attn_out_row(p, b, h, q_row) + half * 32;
        uint4 prev[4];
        if (store && p.accumulate) {
#pragma unroll
            for (int i = 0; i < 4; ++i) prev[i] = reinterpret_cast<const uint4*>(o_ptr)[i];
        }
        uint32_t r[32];
        tmem_ld32(o_addr, r);
        tmem_wait_ld();
        if (store) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
                float f[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) f[k] = __uint_as_float(r[i + k]) * inv_l;
                if (p.accumulate) {
                    const uint4 old = prev[i / 8];
                    f[0] = fmaf(p.out_scale, f[0], bf16_lo(old.x)); f[1] = fmaf(p.out_scale, f[1], bf16_hi(old.x));
                    f[2] = fmaf(p.out_scale, f[2], bf16_lo(old.y)); f[3] = fmaf(p.out_scale, f[3], bf16_hi(old.y));
                    f[4] = fmaf(p.out_scale, f[4], bf16_lo(old.z)); f[5] = fmaf(p.out_scale, f[5], bf16_hi(old.z));
                    f[6] = fmaf(p.out_scale, f[6], bf16_lo(old.w)); f[7] = fmaf(p.out_scale, f[7], bf16_hi(old.w));
                }
                uint4 v;
                v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]);
                v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
                reinterpret_cast<uint4*>(o_ptr + i)[0] = v;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}


#endif  // TG_DEVELOPER (v1 / v2)

// =====================================================================================================================
// v3: v2's tiling and warp roles, with the three things the v2 timeline (tools/attn_trace.py) showed were missing:
//   * the two query tiles TAKE TURNS on the MUFU pipe, per SM sub-partition: the two warps of tile t on sub-partition q
//     start their exponentials only after the two warps of tile 1-t on the same sub-partition have issued theirs
//     (mbarrier hand-off `turn(t, q)`, 2 arrivals).  Left alone, both tiles drift into the same phase: they share the
//     pipe for ~2000 cycles and then both sit in their latency-bound phase (S load, row max, exchange, P store) with the
//     pipe idle.  Alternating, one tile's latency phase hides under the other tile's exponentials.
//   * P is packed and stored to TMEM in 16-column chunks inside the exponential loop, so only the last chunk's store is
//     left on the tail of the block.
//   * a lean instruction stream: hinted mbarrier waits (a blocked warp sleeps instead of spinning through its
//     sub-partition's issue slots), 112 registers for the softmax warps (no address rematerialisation), and the optional
//     FMA-pipe exponentials in packed f32x2 form with the range reduction folded into the score scaling:
//     the reference max is kept as an INTEGER number of log2 units (mc = ceil(max * c)), so
//         t = fma.rm(s, c, 1.5*2^23 - mc)   has floor(s*c - mc) in its low mantissa bits, exactly,
//         f = fma.rn(s, c, -(t - (1.5*2^23 - mc)))  is the fractional part,
//     followed by a degree-3 polynomial and one shift-add into the exponent field (max rel. error 8.8e-5).
#ifndef A3_PRODUCER_REGS
#define A3_PRODUCER_REGS 32  // (96 - producer) * 128 registers are all the softmax warps can gain: 40 -> 104, 32 -> 112
#define A3_SOFTMAX_REGS 112
#endif
#ifndef A3_OPAQUE
#define A3_OPAQUE 1
#endif
#ifndef A3_TURN_AT
#define A3_TURN_AT 4  // 16-column chunks of exponentials issued before the MUFU pipe is handed to the other tile
#endif
constexpr int A3_THREADS = 640;
constexpr int A3_XCH_BYTES = 2 * 2 * 2 * 128 * 4;
constexpr int A3_SMEM_BYTES = 2 * AT_TILE_BYTES + AT_STAGES * 2 * AT_TILE_BYTES + A3_XCH_BYTES + 384 + 1024;

template <int EMU8>
__device__ __forceinline__ constexpr bool a3_emulated(int pair) {  // EMU8 of every 8 pairs, evenly spread
    return ((pair + 1) * EMU8) / 8 != (pair * EMU8) / 8;
}

template <int EMU8, bool ALT>
__global__ void __launch_bounds__(A3_THREADS, 1)
attn3_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                 const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_q2,
                 const __grid_constant__ CUtensorMap tmap_k2, const __grid_constant__ CUtensorMap tmap_v2,
                 const __grid_constant__ AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t q_smem = smem_base;
    const uint32_t kv_smem = smem_base + 2 * AT_TILE_BYTES;
    const uint32_t xch_smem = kv_smem + AT_STAGES * 2 * AT_TILE_BYTES;
    const uint32_t bar_base = xch_smem + A3_XCH_BYTES;
    const uint32_t q_full = bar_base;
    auto kv_full = [&](int s) { return bar_base + 8u * (1 + s); };
    auto kv_empty = [&](int s) { return bar_base + 8u * (1 + AT_STAGES + s); };
    auto s_full = [&](int t) { return bar_base + 8u * (1 + 2 * AT_STAGES + t); };
    auto p_full = [&](int t) { return bar_base + 8u * (3 + 2 * AT_STAGES + t); };
    auto o_done = [&](int t) { return bar_base + 8u * (5 + 2 * AT_STAGES + t); };
    auto s_free = [&](int t) { return bar_base + 8u * (7 + 2 * AT_STAGES + t); };
    auto turn = [&](int t, int q) { return bar_base + 8u * (9 + 2 * AT_STAGES + t * 4 + q); };
    const uint32_t q_empty = bar_base + 8u * (17 + 2 * AT_STAGES);  // the pass's last QK^T has read Q: the next pass may load its Q
    const uint32_t tmem_slot = bar_base + 8u * (18 + 2 * AT_STAGES);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    // Speculative reference (p.spec): the row max of key block 0 stays the softmax reference for the whole pass — no per-block
    // row max, no exchange between the two threads of a row, no rescale.  Exact as long as no exponential leaves the fp32 /
    // bf16 range (powers of two scale exactly): a row whose sum reaches 2^100 (or is not finite) sets `redo_flag`, the CTA
    // stores nothing, and when the regular passes are over every role runs them AGAIN with the exact running-max path
    // (`verdict`: one arrival per softmax warp; producer and MMA issuers wait for it at the end of their last pass, when they
    // would be idle anyway).  With LayerNormed q / k (|s| <= |q||k|/8) the redo never runs; it is what keeps the result exact.
    const uint32_t verdict = bar_base + 8u * (19 + 2 * AT_STAGES);
    const uint32_t redo_flag = bar_base + 8u * (20 + 2 * AT_STAGES);
    volatile uint32_t* redo_flag_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (redo_flag - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int bh = blockIdx.y;
    const int q0 = blockIdx.x * (2 * AT_BLOCK_Q);
    // Up to two passes over the SAME query rows in one CTA (tg_attn_fwd_pair): pass 0 = (Q, K, V), pass 1 = (Q2, K2, V2)
    // accumulated into the same output rows — the self-attention and the text/video -> vip cross-attention of the
    // video-IP-adapter processor.  TMEM, barriers and the K/V ring are set up once; barrier parities run on the cumulative
    // block index g.
    const int n_pass = p.n_pass;
    const int total_pass = p.spec ? 2 * n_pass : n_pass;  // pass indices >= n_pass are the exact redo
    auto pass_blocks = [&](int pass) { return ((pass == 0 ? p.kv_rows : p.kv_rows2) + AT_BLOCK_KV - 1) / AT_BLOCK_KV; };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_k);
        tma_prefetch_desc(&tmap_v);
        if (n_pass > 1) {
            tma_prefetch_desc(&tmap_q2);
            tma_prefetch_desc(&tmap_k2);
            tma_prefetch_desc(&tmap_v2);
        }
    }
    if (warp == 1 && lane == 0) {
        mbar_init(q_full, 1);
        mbar_init(q_empty, 2);
        mbar_init(verdict, 16);
        *redo_flag_ptr = 0u;
        for (int s = 0; s < AT_STAGES; ++s) {
            mbar_init(kv_full(s), 1);
            mbar_init(kv_empty(s), 2);
        }
        for (int t = 0; t < 2; ++t) {
            mbar_init(s_full(t), 1);
            mbar_init(p_full(t), 8);
            mbar_init(s_free(t), 8);
            mbar_init(o_done(t), 1);
            for (int q = 0; q < 4; ++q) mbar_init(turn(t, q), 2);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    auto s_col = [&](int t) { return uint32_t(t * 128); };
    auto o_col = [&](int t) { return uint32_t(256 + t * 64); };
    auto p_col = [&](int t) { return uint32_t(384 + t * 64); };

    if (warp < 4) {
        reg_dealloc<A3_PRODUCER_REGS>();
        if (warp == 0) {
            // -------------------------------------------------------------- TMA producer
            if (lane == 0) {
                int stage = 0;
                uint32_t phase = 0;
                for (int pi = 0; pi < total_pass; ++pi) {
                    if (pi == n_pass) {
                        mbar_wait_fast(verdict, 0u);
                        if (*redo_flag_ptr == 0u) break;
                    }
                    const int pass = pi >= n_pass ? pi - n_pass : pi;
                    const CUtensorMap* mq = pass == 0 ? &tmap_q : &tmap_q2;
                    const CUtensorMap* mk = pass == 0 ? &tmap_k : &tmap_k2;
                    const CUtensorMap* mv = pass == 0 ? &tmap_v : &tmap_v2;
                    if (pi > 0) mbar_wait_fast(q_empty, uint32_t((pi - 1) & 1));
                    mbar_arrive_expect_tx(q_full, 2 * AT_TILE_BYTES);
                    tma_load_3d(q_smem, mq, q_full, 0, q0, bh);
                    tma_load_3d(q_smem + AT_TILE_BYTES, mq, q_full, 0, q0 + AT_BLOCK_Q, bh);
                    const int nb = pass_blocks(pass);
                    for (int j = 0; j < nb; ++j) {
                        mbar_wait_fast(kv_empty(stage), phase ^ 1u);
                        const uint32_t ks = kv_smem + stage * 2 * AT_TILE_BYTES;
                        mbar_arrive_expect_tx(kv_full(stage), 2 * AT_TILE_BYTES);
                        tma_load_3d(ks, mk, kv_full(stage), 0, j * AT_BLOCK_KV, bh);
                        tma_load_3d(ks + AT_TILE_BYTES, mv, kv_full(stage), 0, j * AT_BLOCK_KV, bh);
                        if (++stage == AT_STAGES) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
        } else if (warp == 1 || warp == 3) {
            // -------------------------------------------------------------- MMA issuer of tile t
            if (lane == 0) {
                const int t = warp >> 1;
                constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, false, false);
                constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, false, true);  // V is MN-major
                const uint32_t qs = q_smem + t * AT_TILE_BYTES;
                const uint32_t s_t = tmem_base + s_col(t), o_t = tmem_base + o_col(t), p_t = tmem_base + p_col(t);
                const uint32_t bar_sfull = s_full(t), bar_sfree = s_free(t), bar_pfull = p_full(t), bar_odone = o_done(t);
                auto issue_qk = [&](int stage) {
                    const uint32_t ks = kv_smem + stage * 2 * AT_TILE_BYTES;
#pragma unroll
                    for (int k = 0; k < AT_D / 16; ++k) {
                        const uint64_t da = make_smem_desc_sw128(qs + k * 32, 16, 1024);
                        const uint64_t db = make_smem_desc_sw128(ks + k * 32, 16, 1024);
                        umma_ss(s_t, da, db, idesc_qk, k != 0 ? 1u : 0u);
                    }
                    umma_commit(bar_sfull);
                };
                auto issue_pv = [&](int stage, bool first) {
                    const uint32_t vs = kv_smem + stage * 2 * AT_TILE_BYTES + AT_TILE_BYTES;
#pragma unroll
                    for (int k = 0; k < AT_BLOCK_KV / 16; ++k) {
                        const uint64_t db = make_smem_desc_sw128(vs + k * 2048, 1024, 1024);
                        umma_ts(o_t, p_t + uint32_t(k * 8), db, idesc_pv, (first && k == 0) ? 0u : 1u);
                    }
                };
                int stage = 0;
                uint32_t phase = 0;
                int g = 0;  // cumulative block index over the passes
                for (int pi = 0; pi < total_pass; ++pi) {
                    if (pi == n_pass) {
                        mbar_wait_fast(verdict, 0u);
                        if (*redo_flag_ptr == 0u) break;
                    }
                    const int pass = pi >= n_pass ? pi - n_pass : pi;
                    const int nb = pass_blocks(pass);
                    mbar_wait_fast(q_full, uint32_t(pi & 1));
                    if (g > 0) mbar_wait_fast(bar_sfree, uint32_t((g - 1) & 1));  // the previous pass's last S_t is in registers
                    mbar_wait_fast(kv_full(stage), phase);
                    tc_fence_after();
                    issue_qk(stage);
                    if (nb == 1) umma_commit(q_empty);
                    for (int j = 0; j < nb; ++j, ++g) {
                        int nstage = stage + 1;
                        uint32_t nphase = phase;
                        if (nstage == AT_STAGES) {
                            nstage = 0;
                            nphase ^= 1u;
                        }
                        if (j + 1 < nb) {
                            mbar_wait_fast(bar_sfree, uint32_t(g & 1));  // S_t(j) is in registers: overwrite it
                            mbar_wait_fast(kv_full(nstage), nphase);
                            tc_fence_after();
                            issue_qk(nstage);
                            if (j + 2 == nb) umma_commit(q_empty);  // the pass's last QK^T is in flight: Q may be replaced when it lands
                        }
                        mbar_wait_fast(bar_pfull, uint32_t(g & 1));
                        tc_fence_after();
                        issue_pv(stage, j == 0);
                        umma_commit(kv_empty(stage));
                        umma_commit(bar_odone);
                        stage = nstage;
                        phase = nphase;
                    }
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ softmax warps
        reg_alloc<A3_SOFTMAX_REGS>();
        const int sw = warp - 4;
        const int t = sw >> 3;           // tile
        const int half = (sw >> 2) & 1;  // which 64 of the block's 128 score columns
        const int wq = warp & 3;         // TMEM lane quarter == SM sub-partition
        const int row_in_tile = wq * 32 + lane;
        const int q_row = q0 + t * AT_BLOCK_Q + row_in_tile;
        const uint32_t lane_base = tmem_base + (uint32_t(wq * 32) << 16);
#if A3_OPAQUE
        // loop-invariant addresses pinned in registers (ptxas otherwise re-derives them from %tid every block)
#define A3_PIN(x) asm volatile("mov.b32 %0, %0;" : "+r"(x))
#else
#define A3_PIN(x)
#endif
        uint32_t s_addr = lane_base + s_col(t) + uint32_t(half * 64);
        uint32_t o_addr = lane_base + o_col(t) + uint32_t(half * 32);
        uint32_t p_addr = lane_base + p_col(t) + uint32_t(half * 32);
        uint32_t bar_sfull = s_full(t), bar_sfree = s_free(t), bar_pfull = p_full(t), bar_odone = o_done(t);
        uint32_t bar_my_turn = turn(t, wq), bar_other_turn = turn(1 - t, wq);
        int pair_bar = 1 + t * 4 + wq;
        uint32_t xs_mine = xch_smem + uint32_t(((t * 2 + half) * 128 + row_in_tile) * 4);        // + parity * 2048
        uint32_t xs_other = xch_smem + uint32_t(((t * 2 + (half ^ 1)) * 128 + row_in_tile) * 4);
        uint32_t lane_pin = uint32_t(lane);   // `lane == 0` re-derived from %tid costs S2R + LOP3 at every arrive
        A3_PIN(s_addr); A3_PIN(o_addr); A3_PIN(p_addr); A3_PIN(bar_sfull); A3_PIN(bar_sfree); A3_PIN(bar_pfull);
        A3_PIN(bar_odone); A3_PIN(pair_bar); A3_PIN(xs_mine); A3_PIN(xs_other); A3_PIN(lane_pin);
        auto sts_f32 = [](uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); };
        auto lds_f32 = [](uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory"); return v; };
        const float c = p.scale_log2;
        const float inv_c = 1.0f / c;
        const uint64_t c2 = pack_f32x2(c, c);
        constexpr float MAGIC = 12582912.0f;  // 1.5 * 2^23
        int g = 0;   // cumulative block index over the passes (mbarrier parities)
        int xg = 0;  // exchange-slot parity: advances with every use of the pair's shared-memory slots
        for (int pi = 0; pi < total_pass; ++pi) {
        if (pi == n_pass && *redo_flag_ptr == 0u) break;   // every softmax thread read the final flag after the last vote
        const int pass = pi >= n_pass ? pi - n_pass : pi;
        const bool exact = p.spec == 0 || pi >= n_pass;
        int n_blocks = pass_blocks(pass);
        const int kv_rows = pass == 0 ? p.kv_rows : p.kv_rows2;
        int valid = kv_rows - half * 64;   // score columns of this thread's half that are real keys, from block j on
        A3_PIN(n_blocks); A3_PIN(valid);
        float mc = 0.f;    // reference max in log2 units, integer-valued
        float smin = 0.f;  // scores below this are clamped before an emulated exponential (2^-126)
        float l = 0.f;

        for (int j = 0; j < n_blocks; ++j, ++g) {
            mbar_wait_fast(bar_sfull, uint32_t(g & 1));
            tc_fence_after();
            uint32_t r[64];
            {
                uint32_t (&ra)[32] = *reinterpret_cast<uint32_t (*)[32]>(&r[0]);
                uint32_t (&rb)[32] = *reinterpret_cast<uint32_t (*)[32]>(&r[32]);
                tmem_ld32(s_addr, ra);
                tmem_ld32(s_addr + 32, rb);
                tmem_wait_ld();
            }
            tc_fence_before();
            __syncwarp();
            if (lane_pin == 0) mbar_arrive(bar_sfree);
            if (valid < 64) {
#pragma unroll
                for (int i = 0; i < 64; ++i)
                    if (i >= valid) r[i] = 0xff800000u;  // -inf
            }
            valid -= AT_BLOCK_KV;
#ifdef A3_FIXED_TEST  // developer experiment: upper bound of what an a-priori row bound (no running max) would buy
            bool waited = false;
            if (j == 0) { mc = float(A3_FIXED_TEST); smin = (mc - 126.0f) * inv_c; }
#else
            bool waited = false;
            if (exact || j == 0) {
            float pm[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) pm[k] = fmaxf(__uint_as_float(r[2 * k]), __uint_as_float(r[2 * k + 1]));
#pragma unroll
            for (int i = 8; i < 64; i += 8)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    pm[k] = fmaxf(pm[k], fmaxf(__uint_as_float(r[i + 2 * k]), __uint_as_float(r[i + 2 * k + 1])));
            float mx = fmaxf(fmaxf(pm[0], pm[1]), fmaxf(pm[2], pm[3]));
            {
                const uint32_t par = uint32_t(xg++ & 1) * 2048u;
                sts_f32(xs_mine + par, mx);
                named_bar_sync(pair_bar, 64);
                mx = fmaxf(mx, lds_f32(xs_other + par));
            }
            if (j == 0) {
                mc = ceilf(mx * c);
                smin = (mc - 126.0f) * inv_c;
            } else {
                const bool need = fmaf(mx, c, -mc) > 8.0f;
                if (__any_sync(0xffffffffu, need)) {
                    mbar_wait_fast(bar_odone, uint32_t((g - 1) & 1));
                    tc_fence_after();
                    waited = true;
                    float alpha = 1.0f;
                    if (need) {
                        const float mc_new = ceilf(mx * c);
                        alpha = fast_exp2(mc - mc_new);
                        mc = mc_new;
                        smin = (mc - 126.0f) * inv_c;
                        l *= alpha;
                    }
                    uint32_t ro[32];
                    tmem_ld32(o_addr, ro);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) ro[i] = __float_as_uint(__uint_as_float(ro[i]) * alpha);
                    tmem_st32(o_addr, ro);
                }
            }
            }  // exact || j == 0
#endif
            const uint64_t nmc2 = pack_f32x2(-mc, -mc);
            const float Kf = MAGIC - mc;
            const uint64_t K2 = pack_f32x2(Kf, Kf);
            uint64_t ps2[4] = {0ull, 0ull, 0ull, 0ull};
            if (ALT && (g > 0 || t == 1)) mbar_wait_fast(bar_my_turn, uint32_t((t == 1 ? g : g - 1) & 1));
            // exponentials: a pure FFMA2 / MUFU.EX2 / FADD2 stream (nothing in it waits on a MUFU result except the row sum)
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int i = ch * 16 + 2 * q;
                    float e0, e1;
                    if (a3_emulated<EMU8>(q)) {
                        const float s0 = fmaxf(__uint_as_float(r[i]), smin), s1 = fmaxf(__uint_as_float(r[i + 1]), smin);
                        const uint64_t s2 = pack_f32x2(s0, s1);
                        const uint64_t t2 = fma_rm_f32x2(s2, c2, K2);              // MAGIC + floor(s*c - mc)
                        const uint64_t f2 = fma_f32x2(s2, c2, sub_f32x2(K2, t2));  // fractional part, [0, 1)
                        uint64_t p2 = fma_f32x2(f2, pack_f32x2(0.077119089663028717f, 0.077119089663028717f),
                                                pack_f32x2(0.227564394474029541f, 0.227564394474029541f));
                        p2 = fma_f32x2(p2, f2, pack_f32x2(0.695146143436431885f, 0.695146143436431885f));
                        p2 = fma_f32x2(p2, f2, pack_f32x2(1.0f, 1.0f));
                        e0 = __int_as_float(int(uint32_t(p2)) + (int(uint32_t(t2)) << 23));
                        e1 = __int_as_float(int(uint32_t(p2 >> 32)) + (int(uint32_t(t2 >> 32)) << 23));
                    } else {
                        const uint64_t x2 = fma_f32x2(pack_f32x2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])), c2, nmc2);
                        e0 = fast_exp2(f32x2_lo(x2));
                        e1 = fast_exp2(f32x2_hi(x2));
                    }
                    r[i] = __float_as_uint(e0);
                    r[i + 1] = __float_as_uint(e1);
                }
                if (ALT && ch == A3_TURN_AT - 1) {  // hand the MUFU pipe to the other tile (its wake-up overlaps our tail)
                    asm volatile("" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_other_turn);
                }
            }
#pragma unroll
            for (int i = 0; i < 64; i += 2)
                ps2[(i >> 1) & 3] = add_f32x2(ps2[(i >> 1) & 3], pack_f32x2(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
            uint32_t pk[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) pk[i] = pack_bf16x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
            if (j > 0 && !waited) {  // PV_t(j-1) must have read P_t before it is overwritten (long done by now)
                mbar_wait_fast(bar_odone, uint32_t((g - 1) & 1));
                tc_fence_after();
            }
            tmem_st32(p_addr, pk);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane_pin == 0) mbar_arrive(bar_pfull);
            const uint64_t a = add_f32x2(add_f32x2(ps2[0], ps2[1]), add_f32x2(ps2[2], ps2[3]));
            l += f32x2_lo(a) + f32x2_hi(a);
        }

        // ---- epilogue: O / l -> global (this thread: 32 of the row's 64 output columns)
        {
            const uint32_t par = uint32_t(xg++ & 1) * 2048u;
            sts_f32(xs_mine + par, l);
            named_bar_sync(pair_bar, 64);
            l += lds_f32(xs_other + par);
        }
        bool skip = false;
        if (!exact) {
            // CTA-wide vote on this pass's speculative result (sticky across the passes of a pair launch)
            if (!(l < 1.2676506e30f)) *redo_flag_ptr = 1u;  // 2^100; also catches inf / NaN
            named_bar_sync(9, 512);
            skip = *redo_flag_ptr != 0u;
            if (pi == n_pass - 1) {
                __syncwarp();
                if (lane == 0) mbar_arrive(verdict);
            }
        }
        mbar_wait_fast(bar_odone, uint32_t((g - 1) & 1));
        tc_fence_after();
        const float inv_l = 1.0f / l;
        const bool accumulate = pass == 0 ? (p.accumulate != 0) : true;
        const float out_scale = pass == 0 ? p.out_scale : p.out_scale2;
        const bool store = q_row < p.q_rows && !skip;
        const int b = bh / p.H, h = bh - b * p.H;
        __nv_bfloat16* o_ptr =
            attn_out_row(p, b, h, q_row) + half * 32;
        uint4 prev[4];
        if (store && accumulate) {
#pragma unroll
            for (int i = 0; i < 4; ++i) prev[i] = reinterpret_cast<const uint4*>(o_ptr)[i];
        }
        uint32_t ro[32];
        tmem_ld32(o_addr, ro);
        tmem_wait_ld();
        if (store) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
                float f[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) f[k] = __uint_as_float(ro[i + k]) * inv_l;
                if (accumulate) {
                    const uint4 old = prev[i / 8];
                    f[0] = fmaf(out_scale, f[0], bf16_lo(old.x)); f[1] = fmaf(out_scale, f[1], bf16_hi(old.x));
                    f[2] = fmaf(out_scale, f[2], bf16_lo(old.y)); f[3] = fmaf(out_scale, f[3], bf16_hi(old.y));
                    f[4] = fmaf(out_scale, f[4], bf16_lo(old.z)); f[5] = fmaf(out_scale, f[5], bf16_hi(old.z));
                    f[6] = fmaf(out_scale, f[6], bf16_lo(old.w)); f[7] = fmaf(out_scale, f[7], bf16_hi(old.w));
                }
                uint4 v;
                v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]);
                v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
                reinterpret_cast<uint4*>(o_ptr + i)[0] = v;
            }
        }
        tc_fence_before();  // O_t is in registers: the next pass may overwrite it
        }  // pass
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// Kernel selection.  The shipped library has ONE configuration (constants below: no process-wide mutable state); a developer
// build (-DTG_DEVELOPER, `python -m tokensgen_b200.build --dev`) turns them into knobs behind tg_set_tuning for A/B runs.
#ifdef TG_DEVELOPER
#define TG_KNOB static int
#else
#define TG_KNOB static constexpr int
#endif
TG_KNOB g_attn_impl = 3;  // 1 = v1 (8 softmax warps), 2 = v2 (16), 3 = v3 (16, two tiles per CTA)
TG_KNOB g_attn_emu = 1;   // v3: eighths of the exponentials evaluated on the FMA pipe.  With the speculative reference (fewer issue
                          // slots per score) 1 wins at the power cap too: step 829.5 vs 839.9 ms, 813.6 with the fused pair
                          // launch; 2 loses (825.2) — profiles/r02_ab_bench_2.jsonl
TG_KNOB g_attn_alt = 0;   // v3: the two query tiles take turns on the MUFU pipe
TG_KNOB g_attn_spec = 1;  // v3: speculative softmax reference + exact in-kernel redo (see attn3_fwd_kernel)
#ifdef TG_DEVELOPER
static int g_attn_stagger = 0;
static long long* g_attn_trace = nullptr;
static int g_attn_mutex = 0;
static int g_attn_packed = 1;
#endif

#ifdef TG_DEVELOPER
template <int EMU, bool MUTEX, bool TRACE, bool PACKED>
static int launch_attn2(dim3 grid, cudaStream_t st, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                        const AttnParams& p) {
    static bool attr_set = false;
    auto kern = attn2_fwd_kernel<EMU, MUTEX, TRACE, PACKED>;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, A2_SMEM_BYTES);
        if (e != cudaSuccess) return fail(int(e), "attn_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    kern<<<grid, A2_THREADS, A2_SMEM_BYTES, st>>>(tq, tk, tv, p);
    return check_launch("attn_fwd");
}

#endif

template <int EMU8, bool ALT>
static int launch_attn3(dim3 grid, cudaStream_t st, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                        const CUtensorMap& tq2, const CUtensorMap& tk2, const CUtensorMap& tv2, const AttnParams& p) {
    static bool attr_set = false;
    auto kern = attn3_fwd_kernel<EMU8, ALT>;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, A3_SMEM_BYTES);
        if (e != cudaSuccess) return fail(int(e), "attn_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    kern<<<grid, A3_THREADS, A3_SMEM_BYTES, st>>>(tq, tk, tv, tq2, tk2, tv2, p);
    return check_launch("attn_fwd");
}

static int dispatch_attn3(dim3 grid, cudaStream_t st, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                          const CUtensorMap& tq2, const CUtensorMap& tk2, const CUtensorMap& tv2, const AttnParams& p);

}  // namespace tg

using namespace tg;

// Validates a tg_attn_scatter for output rows [out_row0, out_row0 + q_rows) and copies it into the kernel parameters.
static int set_scatter(AttnParams& p, const tg_attn_scatter* sc, int64_t out_row0, int q_rows, int H) {
    if (sc->world < 1 || sc->world > TG_MAX_PEERS || sc->chunk <= 0 || sc->rows_per_batch <= 0 ||
        int64_t(sc->chunk) * (sc->world - 1) >= sc->rows_per_batch || int64_t(sc->chunk) * sc->world < sc->rows_per_batch)
        return fail(-9, "attn_fwd_sp: bad shard geometry (world=%d chunk=%d rows=%d)", sc->world, sc->chunk, sc->rows_per_batch);
    if (sc->head0 < 0 || sc->head0 + H > sc->H_total) return fail(-9, "attn_fwd_sp: heads [%d, %d) outside H_total=%d", sc->head0, sc->head0 + H, sc->H_total);
    if (out_row0 < 0 || out_row0 + q_rows > sc->rows_per_batch) return fail(-4, "attn_fwd_sp: output rows outside the batch");
    for (int i = 0; i < sc->world; ++i) {
        if (sc->peer[i] == nullptr || (reinterpret_cast<uintptr_t>(sc->peer[i]) & 15))
            return fail(-9, "attn_fwd_sp: no (16-byte aligned) output buffer for rank %d", i);
        p.sp_peer[i] = reinterpret_cast<__nv_bfloat16*>(sc->peer[i]);
    }
    p.sp_world = sc->world;
    p.sp_chunk = sc->chunk;
    p.sp_rows = sc->rows_per_batch;
    p.sp_h_total = sc->H_total;
    p.sp_head0 = sc->head0;
    return 0;
}

static int attn_fwd_impl(const tg_bf16* q, int64_t q_rows_alloc, int64_t q_row0, int q_rows, const tg_bf16* k,
                         const tg_bf16* v, int64_t kv_rows_alloc, int64_t kv_row0, int kv_rows, tg_bf16* out,
                         int64_t out_rows_alloc, int64_t out_row0, int B, int H, float softmax_scale, int accumulate,
                         float out_scale, const tg_attn_scatter* scatter, void* stream) {
    if (scatter != nullptr) { out = scatter->peer[0]; out_rows_alloc = scatter->rows_per_batch; }
    if (!q || !k || !v || !out) return fail(-1, "attn_fwd: null pointer");
    if (B <= 0 || H <= 0 || q_rows <= 0 || kv_rows <= 0) return fail(-2, "attn_fwd: non-positive shape");
    if (q_row0 < 0 || kv_row0 < 0 || q_row0 + q_rows > q_rows_alloc || kv_row0 + kv_rows > kv_rows_alloc)
        return fail(-3, "attn_fwd: row window outside the allocation");
    if (out_row0 < 0 || out_row0 + q_rows > out_rows_alloc) return fail(-4, "attn_fwd: output window outside the allocation");
    if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
         reinterpret_cast<uintptr_t>(out)) & 15)
        return fail(-5, "attn_fwd: pointers must be 16-byte aligned");
    const int BH = B * H;
    if (BH > 65535) return fail(-6, "attn_fwd: B*H too large");
    CUtensorMap tq, tk, tv;
    int rc = make_tmap_3d(&tq, q + q_row0 * AT_D, AT_D, uint64_t(q_rows), uint64_t(BH), AT_D * 2,
                          uint64_t(q_rows_alloc) * AT_D * 2, AT_D, AT_BLOCK_Q);
    if (rc) return rc;
    rc = make_tmap_3d(&tk, k + kv_row0 * AT_D, AT_D, uint64_t(kv_rows), uint64_t(BH), AT_D * 2,
                      uint64_t(kv_rows_alloc) * AT_D * 2, AT_D, AT_BLOCK_KV);
    if (rc) return rc;
    rc = make_tmap_3d(&tv, v + kv_row0 * AT_D, AT_D, uint64_t(kv_rows), uint64_t(BH), AT_D * 2,
                      uint64_t(kv_rows_alloc) * AT_D * 2, AT_D, AT_BLOCK_KV);
    if (rc) return rc;
    AttnParams p{};
    p.q_rows = q_rows;
    p.kv_rows = kv_rows;
    p.H = H;
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    p.out_rows_alloc = out_rows_alloc;
    p.out_row0 = out_row0;
    p.scale_log2 = softmax_scale * 1.4426950408889634f;
    p.accumulate = accumulate;
    p.out_scale = out_scale;
    p.n_pass = 1;
    p.spec = g_attn_spec;
#ifdef TG_DEVELOPER
    p.stagger = g_attn_stagger;
    p.trace = g_attn_trace;
    p.mutex = g_attn_mutex;
#endif
    if (scatter != nullptr) {
        rc = set_scatter(p, scatter, out_row0, q_rows, H);
        if (rc) return rc;
    }
    dim3 grid((q_rows + 2 * AT_BLOCK_Q - 1) / (2 * AT_BLOCK_Q), BH);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#ifdef TG_DEVELOPER
    if (g_attn_impl == 2) {
        if (g_attn_trace != nullptr) return launch_attn2<0, false, true, false>(grid, st, tq, tk, tv, p);
        return g_attn_packed ? launch_attn2<0, false, false, true>(grid, st, tq, tk, tv, p)
                             : launch_attn2<0, false, false, false>(grid, st, tq, tk, tv, p);
    }
    if (g_attn_impl == 1) {
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM_BYTES);
            if (e != cudaSuccess) return fail(int(e), "attn_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            attr_set = true;
        }
        attn_fwd_kernel<<<grid, AT_THREADS, AT_SMEM_BYTES, st>>>(tq, tk, tv, p);
        return check_launch("attn_fwd");
    }
#endif
    return dispatch_attn3(grid, st, tq, tk, tv, tq, tk, tv, p);
}

extern "C" int tg_attn_fwd(const tg_bf16* q, int64_t q_rows_alloc, int64_t q_row0, int q_rows, const tg_bf16* k,
                           const tg_bf16* v, int64_t kv_rows_alloc, int64_t kv_row0, int kv_rows, tg_bf16* out,
                           int64_t out_rows_alloc, int64_t out_row0, int B, int H, float softmax_scale, int accumulate,
                           float out_scale, void* stream) {
    return attn_fwd_impl(q, q_rows_alloc, q_row0, q_rows, k, v, kv_rows_alloc, kv_row0, kv_rows, out, out_rows_alloc, out_row0,
                         B, H, softmax_scale, accumulate, out_scale, nullptr, stream);
}

extern "C" int tg_attn_fwd_sp(const tg_bf16* q, int64_t q_rows_alloc, int64_t q_row0, int q_rows, const tg_bf16* k,
                              const tg_bf16* v, int64_t kv_rows_alloc, int64_t kv_row0, int kv_rows,
                              const tg_attn_scatter* out, int64_t out_row0, int B, int H, float softmax_scale,
                              int accumulate, float out_scale, void* stream) {
    if (out == nullptr) return fail(-1, "attn_fwd_sp: scatter is null");
    return attn_fwd_impl(q, q_rows_alloc, q_row0, q_rows, k, v, kv_rows_alloc, kv_row0, kv_rows, nullptr, 0, out_row0, B, H,
                         softmax_scale, accumulate, out_scale, out, stream);
}

static int tg::dispatch_attn3(dim3 grid, cudaStream_t st, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                              const CUtensorMap& tq2, const CUtensorMap& tk2, const CUtensorMap& tv2, const AttnParams& p) {
#ifdef TG_DEVELOPER
#define TG_A3(E)                                                                                      \
    case E:                                                                                           \
        return g_attn_alt ? launch_attn3<E, true>(grid, st, tq, tk, tv, tq2, tk2, tv2, p)             \
                          : launch_attn3<E, false>(grid, st, tq, tk, tv, tq2, tk2, tv2, p);
    switch (g_attn_emu) {
        TG_A3(0) TG_A3(1) TG_A3(2) TG_A3(3) TG_A3(4)
        default: return fail(-7, "attn_fwd: attn_emu must be 0..4 (eighths of the exponentials on the FMA pipe)");
    }
#undef TG_A3
#else
    return launch_attn3<g_attn_emu, false>(grid, st, tq, tk, tv, tq2, tk2, tv2, p);   // the one shipped instantiation
#endif
}

static int attn_fwd_pair_impl(const tg_bf16* q, const tg_bf16* k, const tg_bf16* v, int64_t rows_alloc, int q_rows, int kv_rows,
                              const tg_bf16* q2, const tg_bf16* k2, const tg_bf16* v2, int64_t rows_alloc2, int64_t kv_row0_2,
                              int kv_rows2, tg_bf16* out, int64_t out_rows_alloc, int B, int H, float softmax_scale,
                              float out_scale2, const tg_attn_scatter* scatter, void* stream) {
    if (scatter != nullptr) { out = scatter->peer[0]; out_rows_alloc = scatter->rows_per_batch; }
    if (!q || !k || !v || !q2 || !k2 || !v2 || !out) return fail(-1, "attn_fwd_pair: null pointer");
    if (B <= 0 || H <= 0 || q_rows <= 0 || kv_rows <= 0 || kv_rows2 <= 0) return fail(-2, "attn_fwd_pair: non-positive shape");
    if (q_rows > rows_alloc || kv_rows > rows_alloc || q_rows > rows_alloc2 || kv_row0_2 < 0 || kv_row0_2 + kv_rows2 > rows_alloc2)
        return fail(-3, "attn_fwd_pair: row window outside the allocation");
    if (q_rows > out_rows_alloc) return fail(-4, "attn_fwd_pair: output window outside the allocation");
    if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
         reinterpret_cast<uintptr_t>(q2) | reinterpret_cast<uintptr_t>(k2) | reinterpret_cast<uintptr_t>(v2) |
         reinterpret_cast<uintptr_t>(out)) & 15)
        return fail(-5, "attn_fwd_pair: pointers must be 16-byte aligned");
    const int BH = B * H;
    if (BH > 65535) return fail(-6, "attn_fwd_pair: B*H too large");
#ifdef TG_DEVELOPER
    if (g_attn_impl != 3) return fail(-8, "attn_fwd_pair: needs attention kernel generation 3");
#endif
    CUtensorMap tq, tk, tv, tq2, tk2, tv2;
    int rc;
    auto mk = [&](CUtensorMap* m, const tg_bf16* base, int64_t row0, int rows, int64_t alloc, int box_rows) {
        return make_tmap_3d(m, base + row0 * AT_D, AT_D, uint64_t(rows), uint64_t(BH), AT_D * 2, uint64_t(alloc) * AT_D * 2, AT_D,
                            box_rows);
    };
    if ((rc = mk(&tq, q, 0, q_rows, rows_alloc, AT_BLOCK_Q))) return rc;
    if ((rc = mk(&tk, k, 0, kv_rows, rows_alloc, AT_BLOCK_KV))) return rc;
    if ((rc = mk(&tv, v, 0, kv_rows, rows_alloc, AT_BLOCK_KV))) return rc;
    if ((rc = mk(&tq2, q2, 0, q_rows, rows_alloc2, AT_BLOCK_Q))) return rc;
    if ((rc = mk(&tk2, k2, kv_row0_2, kv_rows2, rows_alloc2, AT_BLOCK_KV))) return rc;
    if ((rc = mk(&tv2, v2, kv_row0_2, kv_rows2, rows_alloc2, AT_BLOCK_KV))) return rc;
    AttnParams p{};
    p.q_rows = q_rows;
    p.kv_rows = kv_rows;
    p.H = H;
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    p.out_rows_alloc = out_rows_alloc;
    p.out_row0 = 0;
    p.scale_log2 = softmax_scale * 1.4426950408889634f;
    p.accumulate = 0;
    p.out_scale = 1.0f;
    p.n_pass = 2;
    p.spec = g_attn_spec;
    p.kv_rows2 = kv_rows2;
    p.out_scale2 = out_scale2;
    if (scatter != nullptr) {
        rc = set_scatter(p, scatter, 0, q_rows, H);
        if (rc) return rc;
    }
    dim3 grid((q_rows + 2 * AT_BLOCK_Q - 1) / (2 * AT_BLOCK_Q), BH);
    return dispatch_attn3(grid, static_cast<cudaStream_t>(stream), tq, tk, tv, tq2, tk2, tv2, p);
}

extern "C" int tg_attn_fwd_pair(const tg_bf16* q, const tg_bf16* k, const tg_bf16* v, int64_t rows_alloc, int q_rows, int kv_rows,
                                const tg_bf16* q2, const tg_bf16* k2, const tg_bf16* v2, int64_t rows_alloc2, int64_t kv_row0_2,
                                int kv_rows2, tg_bf16* out, int64_t out_rows_alloc, int B, int H, float softmax_scale,
                                float out_scale2, void* stream) {
    return attn_fwd_pair_impl(q, k, v, rows_alloc, q_rows, kv_rows, q2, k2, v2, rows_alloc2, kv_row0_2, kv_rows2, out,
                              out_rows_alloc, B, H, softmax_scale, out_scale2, nullptr, stream);
}

extern "C" int tg_attn_fwd_pair_sp(const tg_bf16* q, const tg_bf16* k, const tg_bf16* v, int64_t rows_alloc, int q_rows,
                                   int kv_rows, const tg_bf16* q2, const tg_bf16* k2, const tg_bf16* v2, int64_t rows_alloc2,
                                   int64_t kv_row0_2, int kv_rows2, const tg_attn_scatter* out, int B, int H,
                                   float softmax_scale, float out_scale2, void* stream) {
    if (out == nullptr) return fail(-1, "attn_fwd_pair_sp: scatter is null");
    return attn_fwd_pair_impl(q, k, v, rows_alloc, q_rows, kv_rows, q2, k2, v2, rows_alloc2, kv_row0_2, kv_rows2, nullptr, 0, B, H,
                              softmax_scale, out_scale2, out, stream);
}

#ifdef TG_DEVELOPER
// Developer hooks (tools/attn_*.py, tools/ab_bench.sh): present only in libtokensgen_b200_dev.so, never in the shipped library.
extern "C" int tg_debug_attn_trace(void* device_buffer) {  // developer hook, not in the public header
    g_attn_trace = static_cast<long long*>(device_buffer);
    return 0;
}

extern "C" int tg_set_tuning(const char* key, int value) {
    if (key == nullptr) return fail(-1, "set_tuning: null key");
    const std::string k(key);
    if (k == "attn_impl") { g_attn_impl = value; return 0; }
    if (k == "attn_emu") { g_attn_emu = value; return 0; }
    if (k == "attn_stagger") { g_attn_stagger = value; return 0; }
    if (k == "attn_mutex") { g_attn_mutex = value; return 0; }
    if (k == "attn_packed") { g_attn_packed = value; return 0; }
    if (k == "attn_alt") { g_attn_alt = value; return 0; }
    if (k == "attn_spec") { g_attn_spec = value; return 0; }
    return fail(-2, "set_tuning: unknown key %s", key);
}
#endif  // TG_DEVELOPER
