// tcgen05 GEMM family for the DiT projections.
//
//   out = epilogue(A[M,K] @ W[N,K]^T + bias)      A, W bf16 K-major; fp32 accumulation in TMEM.
//
// One persistent CTA per SM, 256 threads:
//   warp 0      TMA producer   (A tile 128x64, W tile BLOCK_Nx64 per stage; 128-byte swizzle)
//   warp 1      MMA issuer     (one elected lane: 4 x tcgen05.mma 128xBLOCK_Nx16 per stage)
//   warp 2      TMEM allocator (2 accumulator buffers of BLOCK_N columns: epilogue of tile i overlaps mainloop of i+1)
//   warps 4..7  epilogue       (thread = one output row: tcgen05.ld -> fused epilogue -> global)
// Pipelines: smem full/empty mbarriers (TMA <-> MMA), tmem full/empty mbarriers (MMA <-> epilogue).
//
// Epilogues (all row-local because one thread owns one accumulator row):
//   EPI_BIAS_ACT       bias (+ GELU-tanh / SiLU), bf16 store
//   EPI_GATE_RESIDUAL  X[m,:] += gate(m) * (acc + bias), in place           (to_out / ff.net.2 + AdaLN-Zero gate)
//   EPI_QKV            bias, per-head LayerNorm(64), 3D-RoPE, head-major store [B,H,rows,64]
#include <cuda_bf16.h>

#include "common.h"
#include "ptx.cuh"

namespace tg {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int GROUP_M = 16;  // m-tiles per scheduling group (L2 reuse of the weight tile)

enum { EPI_BIAS_ACT = 0, EPI_GATE_RESIDUAL = 1, EPI_QKV = 2, EPI_QKV_SP = 3 };  // _SP: heads scattered to peer ranks
__host__ __device__ constexpr bool epi_is_qkv(int e) { return e == EPI_QKV || e == EPI_QKV_SP; }

struct QkvProjDev {
    __nv_bfloat16* out;
    int out_rows;
    const __nv_bfloat16* ln_w;
    const __nv_bfloat16* ln_b;
    const float* cos_video;
    const float* sin_video;
    const float* cos_vip;
    const float* sin_vip;
};

struct GemmParams {
    int M, N, K;
    int m_tiles, n_tiles;
    const __nv_bfloat16* bias;  // [N] or null
    __nv_bfloat16* out;         // EPI_BIAS_ACT: out; EPI_GATE_RESIDUAL: X (in place)
    int64_t ldo;
    int act;
    tg_rowmap map;
    tg_modvec gate;
    // EPI_QKV
    QkvProjDev proj[6];
    int heads;
    float ln_eps;
    // EPI_QKV, sequence-parallel scatter (tg_qkv_rope_gemm_sp): sp_world > 1 -> head h of projection j goes to
    // sp_peer[j][h / sp_hpr], laid out [B, sp_hpr, out_rows, 64]
    int sp_world, sp_hpr;
    __nv_bfloat16* sp_peer[6][TG_MAX_PEERS];
};

template <int BLOCK_N>
struct GemmCfg {
    static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
    static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
    static constexpr int TMEM_COLS = 2 * BLOCK_N;  // 512 / 256 / 128
    static constexpr int BAR_BYTES = 256;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024: manual 1 KB alignment
};

struct RowInfo {
    int b, r, seg, frame;  // seg: 0 text, 1 video, 2 vip
};
__device__ __forceinline__ RowInfo row_info(const tg_rowmap& m, int row) {
    RowInfo ri;
    ri.b = row / m.rows_local;  // host entries pass normalised_rowmap(): rows_local > 0
    ri.r = row - ri.b * m.rows_local + m.row0;
    if (ri.r < m.n_text) {
        ri.seg = 0;
        ri.frame = 0;
    } else if (ri.r < m.n_text + m.n_video) {
        ri.seg = 1;
        ri.frame = (ri.r - m.n_text) / m.hw;
    } else {
        ri.seg = 2;
        ri.frame = 0;
    }
    return ri;
}
__device__ __forceinline__ const __nv_bfloat16* modvec_row(const tg_modvec& v, const tg_rowmap& m, const RowInfo& ri) {
    const tg_bf16* p;
    if (ri.seg == 0)
        p = v.text ? v.text + int64_t(ri.b * m.frames) * v.ld_text : nullptr;
    else if (ri.seg == 1)
        p = v.video ? v.video + int64_t(ri.b * m.frames + ri.frame) * v.ld_video : nullptr;
    else
        p = v.vip ? v.vip + int64_t(ri.b * m.frames) * v.ld_vip : nullptr;
    return reinterpret_cast<const __nv_bfloat16*>(p);
}

__device__ __forceinline__ void tile_coords(int tile, int m_tiles, int n_tiles, int& mt, int& nt) {
    const int group_sz = GROUP_M * n_tiles;
    const int g = tile / group_sz;
    const int first_m = g * GROUP_M;
    const int gm = min(GROUP_M, m_tiles - first_m);
    const int in_g = tile - g * group_sz;
    mt = first_m + in_g % gm;
    nt = in_g / gm;
}

__device__ __forceinline__ void load8_bf16(const __nv_bfloat16* p, float (&f)[8]) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));  // read-only operands only (bias, norm weights)
    f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
    f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ void store8_bf16(__nv_bfloat16* p, const float* f) {
    uint4 v;
    v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]);
    v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = v;
}

// ---- epilogue bodies: `acc` = 32 or 64 fp32 accumulator columns of one row, starting at global column n.
template <int NC>
__device__ __forceinline__ void add_bias(float (&acc)[NC], const __nv_bfloat16* bias, int n) {
    if (bias == nullptr) return;
#pragma unroll
    for (int i = 0; i < NC; i += 8) {
        float b[8];
        load8_bf16(bias + n + i, b);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i + j] += b[j];
    }
}

__device__ __forceinline__ void epi_bias_act(const GemmParams& p, float (&acc)[32], int row, int n) {
    add_bias<32>(acc, p.bias, n);
    if (p.act == TG_ACT_GELU_TANH) {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = gelu_tanh(acc[i]);
    } else if (p.act == TG_ACT_SILU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = __fdividef(acc[i], 1.0f + __expf(-acc[i]));
    }
    if (row < p.M) {
        __nv_bfloat16* o = p.out + int64_t(row) * p.ldo + n;
#pragma unroll
        for (int i = 0; i < 32; i += 8) store8_bf16(o + i, &acc[i]);
    }
}

__device__ __forceinline__ void epi_gate_residual(const GemmParams& p, float (&acc)[32], int row, int n) {
    add_bias<32>(acc, p.bias, n);
    if (row < p.M) {
        const RowInfo ri = row_info(p.map, row);
        const uint4* g = reinterpret_cast<const uint4*>(modvec_row(p.gate, p.map, ri) + n);
        uint4* x = reinterpret_cast<uint4*>(p.out + int64_t(row) * p.ldo + n);
        // all loads first, then all stores: a store to X followed by a load of the neighbouring 16 bytes of the same line
        // would serialise on an L2 round trip per 8 elements
        uint4 xr[4], gr[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) xr[i] = x[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) gr[i] = __ldg(g + i);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float gv[8], xv[8];
            gv[0] = bf16_lo(gr[i].x); gv[1] = bf16_hi(gr[i].x); gv[2] = bf16_lo(gr[i].y); gv[3] = bf16_hi(gr[i].y);
            gv[4] = bf16_lo(gr[i].z); gv[5] = bf16_hi(gr[i].z); gv[6] = bf16_lo(gr[i].w); gv[7] = bf16_hi(gr[i].w);
            xv[0] = bf16_lo(xr[i].x); xv[1] = bf16_hi(xr[i].x); xv[2] = bf16_lo(xr[i].y); xv[3] = bf16_hi(xr[i].y);
            xv[4] = bf16_lo(xr[i].z); xv[5] = bf16_hi(xr[i].z); xv[6] = bf16_lo(xr[i].w); xv[7] = bf16_hi(xr[i].w);
#pragma unroll
            for (int j = 0; j < 8; ++j) xv[j] = fmaf(gv[j], acc[i * 8 + j], xv[j]);
            xr[i].x = pack_bf16x2(xv[0], xv[1]); xr[i].y = pack_bf16x2(xv[2], xv[3]);
            xr[i].z = pack_bf16x2(xv[4], xv[5]); xr[i].w = pack_bf16x2(xv[6], xv[7]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = xr[i];
    }
}

// One head (64 columns) of one projection for one row: bias, LayerNorm(64), RoPE in place on `acc`; returns where the
// row's 64 bf16 values go (nullptr: the row is not stored).
template <bool SP>
__device__ __forceinline__ __nv_bfloat16* epi_qkv_compute(const GemmParams& p, float (&acc)[64], int row, int n) {
    add_bias<64>(acc, p.bias, n);
    const int inner = p.heads * 64;
    const int pi = n / inner;
    const int head = (n - pi * inner) >> 6;
    const QkvProjDev& pr = p.proj[pi];
    if (pr.ln_w != nullptr) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 64; ++i) s += acc[i];
        const float mean = s * (1.0f / 64.0f);
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 64; ++i) {
            const float d = acc[i] - mean;
            ss = fmaf(d, d, ss);
        }
        const float rstd = rsqrtf(ss * (1.0f / 64.0f) + p.ln_eps);
#pragma unroll
        for (int i = 0; i < 64; i += 8) {
            float w[8], b[8];
            load8_bf16(pr.ln_w + i, w);
            load8_bf16(pr.ln_b + i, b);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i + j] = fmaf((acc[i + j] - mean) * rstd, w[j], b[j]);
        }
    }
    if (row >= p.M) return nullptr;
    const RowInfo ri = row_info(p.map, row);
    if (ri.r >= pr.out_rows) return nullptr;
    const float* cs = nullptr;
    const float* sn = nullptr;
    if (ri.seg == 1 && pr.cos_video != nullptr) {
        cs = pr.cos_video + int64_t(ri.r - p.map.n_text) * 64;
        sn = pr.sin_video + int64_t(ri.r - p.map.n_text) * 64;
    } else if (ri.seg == 2 && pr.cos_vip != nullptr) {
        cs = pr.cos_vip + int64_t(ri.r - p.map.n_text - p.map.n_video) * 64;
        sn = pr.sin_vip + int64_t(ri.r - p.map.n_text - p.map.n_video) * 64;
    }
    if (cs != nullptr) {
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
            const float4 c = *reinterpret_cast<const float4*>(cs + i);
            const float4 s = *reinterpret_cast<const float4*>(sn + i);
            const float x0 = acc[i], x1 = acc[i + 1], x2 = acc[i + 2], x3 = acc[i + 3];
            acc[i] = x0 * c.x - x1 * s.x;
            acc[i + 1] = x1 * c.y + x0 * s.y;
            acc[i + 2] = x2 * c.z - x3 * s.z;
            acc[i + 3] = x3 * c.w + x2 * s.w;
        }
    }
    __nv_bfloat16* o;
    if constexpr (SP) {  // Ulysses all-to-all fused into the store: the head's owner rank receives the row over NVLink
        const int dst = head / p.sp_hpr;
        o = p.sp_peer[pi][dst] + (int64_t(ri.b * p.sp_hpr + (head - dst * p.sp_hpr)) * pr.out_rows + ri.r) * 64;
    } else {
        o = pr.out + (int64_t(ri.b * p.heads + head) * pr.out_rows + ri.r) * 64;
    }
    return o;
}

template <bool SP>
__device__ __forceinline__ void epi_qkv(const GemmParams& p, float (&acc)[64], int row, int n) {
    __nv_bfloat16* o = epi_qkv_compute<SP>(p, acc, row, n);
    if (o == nullptr) return;
#pragma unroll
    for (int i = 0; i < 64; i += 8) store8_bf16(o + i, &acc[i]);
}

// Warp-collective store of 32 rows x 128 bytes through a 4 KB shared-memory transpose: every store instruction then covers
// four full 128-byte lines (8 lanes per row) instead of 32 scattered 16-byte pieces.  Local L2 merges such pieces for
// free; an NVLink peer write does not — each piece is its own packet — and the sequence-parallel Q/K/V scatter ran at
// ~130 GB/s that way.  `o` = this lane's row destination (nullptr = skip).  Chunk slots are XOR-swizzled by the row so
// both the row-wise writes and the line-wise reads are bank-conflict free per quarter warp.
__device__ __forceinline__ void store_rows_coalesced(uint32_t stage, int lane, __nv_bfloat16* o, const float (&acc)[64]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint32_t a = stage + uint32_t(lane) * 128u + (uint32_t(c ^ (lane & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack_bf16x2(acc[c * 8], acc[c * 8 + 1])),
                     "r"(pack_bf16x2(acc[c * 8 + 2], acc[c * 8 + 3])), "r"(pack_bf16x2(acc[c * 8 + 4], acc[c * 8 + 5])),
                     "r"(pack_bf16x2(acc[c * 8 + 6], acc[c * 8 + 7]))
                     : "memory");
    }
    __syncwarp();
    const unsigned long long optr = reinterpret_cast<unsigned long long>(o);
    const int c = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int src = 4 * i + (lane >> 3);
        const unsigned long long po = __shfl_sync(0xffffffffu, optr, src);
        uint4 v;
        const uint32_t a = stage + uint32_t(src) * 128u + (uint32_t(c ^ (src & 7)) << 4);
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
        if (po != 0ull) *reinterpret_cast<uint4*>(po + uint32_t(c) * 16u) = v;
    }
    __syncwarp();  // the staging rows are rewritten by the next head
}

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(256, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const __grid_constant__ GemmParams p) {
    using Cfg = GemmCfg<BLOCK_N>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.m_tiles * p.n_tiles;
    const int k_blocks = p.K / BLOCK_K;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 4);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int mt, nt;
                tile_coords(tile, p.m_tiles, p.n_tiles, mt, nt);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u, 0x101);
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
                    tma_load_2d(sa, &tmap_a, full_bar(stage), kb * BLOCK_K, mt * BLOCK_M);
                    tma_load_2d(sa + Cfg::A_BYTES, &tmap_b, full_bar(stage), kb * BLOCK_K, nt * BLOCK_N);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N, false, false);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u, 0x102);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + uint32_t(acc * BLOCK_N);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(full_bar(stage), phase, 0x103);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        const uint64_t da = make_smem_desc_sw128(sa + k * UMMA_K * 2, 16, 1024);
                        const uint64_t db = make_smem_desc_sw128(sb + k * UMMA_K * 2, 16, 1024);
                        umma_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(empty_bar(stage));  // smem slot free once these MMAs have read it
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit(tfull_bar(acc));  // accumulator complete
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------ epilogue
        const int ew = warp & 3;  // TMEM lane quarter this warp may touch
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int mt, nt;
            tile_coords(tile, p.m_tiles, p.n_tiles, mt, nt);
            const int row = mt * BLOCK_M + ew * 32 + lane;
            mbar_wait(tfull_bar(acc), acc_phase, 0x104);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (uint32_t(ew * 32) << 16) + uint32_t(acc * BLOCK_N);
            if constexpr (epi_is_qkv(EPI)) {
#pragma unroll 1
                for (int c = 0; c < BLOCK_N; c += 64) {
                    uint32_t r0[32], r1[32];
                    tmem_ld32(t_row + c, r0);
                    tmem_ld32(t_row + c + 32, r1);
                    tmem_wait_ld();
                    float accv[64];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        accv[i] = __uint_as_float(r0[i]);
                        accv[32 + i] = __uint_as_float(r1[i]);
                    }
                    epi_qkv<EPI == EPI_QKV_SP>(p, accv, row, nt * BLOCK_N + c);
                }
            } else {
#pragma unroll 1
                for (int c = 0; c < BLOCK_N; c += 32) {
                    uint32_t r0[32];
                    tmem_ld32(t_row + c, r0);
                    tmem_wait_ld();
                    float accv[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) accv[i] = __uint_as_float(r0[i]);
                    if constexpr (EPI == EPI_BIAS_ACT)
                        epi_bias_act(p, accv, row, nt * BLOCK_N + c);
                    else
                        epi_gate_residual(p, accv, row, nt * BLOCK_N + c);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// =====================================================================================================================
// CTA-pair GEMM (cta_group::2): one 2-CTA cluster = one 256 x 256 output tile.  CTA r of the pair loads A rows
// [128 r, 128 r + 128) and W rows (output columns) [128 r, 128 r + 128) of the tile — 32 KB per 64-deep k-block instead of
// the 48 KB a single-CTA 128 x 256 tile needs — and the leader's one MMA thread issues tcgen05.mma.cta_group::2
// (M = 256, N = 256): each SM's tensor core reads its own A half and BOTH W halves (the peer's over the pair link) and
// accumulates its 128 rows x 256 columns in its own TMEM.  Fewer operand bytes per flop through L2 / TMA / shared
// memory is what this buys on a power-capped part.
//   full barrier   (leader only) : leader's expect_tx covers both CTAs' 2 x 32 KB; the peer's TMA completes on it remotely
//   empty barrier  (both CTAs)   : leader's tcgen05.commit multicast -> each CTA's producer refills its own half
//   tmem full      (both CTAs)   : commit multicast -> each CTA's epilogue warps drain their own 128 rows
//   tmem empty     (leader only) : 4 local + 4 remote arrivals (the peer's epilogue warps arrive through the cluster)
constexpr int G2_STAGES = 6;
constexpr int G2_A_BYTES = 128 * BLOCK_K * 2;  // 16 KB
constexpr int G2_B_BYTES = 128 * BLOCK_K * 2;  // 16 KB (this CTA's half of the 256 output columns)
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;
constexpr int G2_XPOSE_BYTES = 4 * 32 * 128;  // EPI_QKV peer scatter: one 32-row x 128-byte transpose buffer per epilogue warp
__host__ __device__ constexpr int g2_smem_bytes(int epi) {
    return G2_STAGES * G2_STAGE_BYTES + 256 + (epi == EPI_QKV_SP ? G2_XPOSE_BYTES : 0) + 1024;
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
             const __grid_constant__ GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + G2_STAGES * G2_STAGE_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (G2_STAGES + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * G2_STAGES + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * G2_STAGES + 2 + a); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * G2_STAGES + 4);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int n_clusters = gridDim.x >> 1;
    const int cluster_id = blockIdx.x >> 1;
    const int num_tiles = p.m_tiles * p.n_tiles;  // 256 x 256 tiles
    const int k_blocks = p.K / BLOCK_K;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < G2_STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 8);  // 4 epilogue warps of each CTA of the pair
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_2cta(tmem_slot, 512);
    tc_fence_before();
    cluster_sync();  // barrier inits and the TMEM allocation of both CTAs are visible pair-wide
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer (both CTAs; bytes land on the LEADER's full barrier)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
                int mt, nt;
                tile_coords(tile, p.m_tiles, p.n_tiles, mt, nt);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u, 0x181);
                    const uint32_t sa = smem_base + stage * G2_STAGE_BYTES;
                    const uint32_t bar = mapa_shared(full_bar(stage), 0);
                    if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * G2_STAGE_BYTES);
                    tma_load_2d_2cta(sa, &tmap_a, bar, kb * BLOCK_K, mt * 256 + int(rank) * 128);
                    tma_load_2d_2cta(sa + G2_A_BYTES, &tmap_b, bar, kb * BLOCK_K, nt * 256 + int(rank) * 128);
                    if (++stage == G2_STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer (leader CTA only)
        if (lane == 0 && leader) {
            constexpr uint32_t idesc = make_idesc_bf16(256, 256, false, false);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u, 0x182);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + uint32_t(acc * 256);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(full_bar(stage), phase, 0x183);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * G2_STAGE_BYTES;
                    const uint32_t sb = sa + G2_A_BYTES;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        const uint64_t da = make_smem_desc_sw128(sa + k * UMMA_K * 2, 16, 1024);
                        const uint64_t db = make_smem_desc_sw128(sb + k * UMMA_K * 2, 16, 1024);
                        umma_ss_2cta(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit_2cta(empty_bar(stage), 0b11);  // both CTAs' smem slots are free once these MMAs have read them
                    if (++stage == G2_STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit_2cta(tfull_bar(acc), 0b11);  // both CTAs' accumulator halves are complete
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------ epilogue (both CTAs, each its own 128 rows x 256 columns)
        const int ew = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
            int mt, nt;
            tile_coords(tile, p.m_tiles, p.n_tiles, mt, nt);
            const int row = mt * 256 + int(rank) * 128 + ew * 32 + lane;
            mbar_wait(tfull_bar(acc), acc_phase, 0x184);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (uint32_t(ew * 32) << 16) + uint32_t(acc * 256);
            if constexpr (epi_is_qkv(EPI)) {
#pragma unroll 1
                for (int c = 0; c < 256; c += 64) {
                    uint32_t r0[32], r1[32];
                    tmem_ld32(t_row + c, r0);
                    tmem_ld32(t_row + c + 32, r1);
                    tmem_wait_ld();
                    float accv[64];
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        accv[i] = __uint_as_float(r0[i]);
                        accv[32 + i] = __uint_as_float(r1[i]);
                    }
                    if constexpr (EPI == EPI_QKV_SP) {
                        __nv_bfloat16* o = epi_qkv_compute<true>(p, accv, row, nt * 256 + c);
                        store_rows_coalesced(bar_base + 256u + uint32_t(ew) * 4096u, lane, o, accv);
                    } else {
                        epi_qkv<false>(p, accv, row, nt * 256 + c);
                    }
                }
            } else {
#pragma unroll 1
                for (int c = 0; c < 256; c += 32) {
                    uint32_t r0[32];
                    tmem_ld32(t_row + c, r0);
                    tmem_wait_ld();
                    float accv[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) accv[i] = __uint_as_float(r0[i]);
                    if constexpr (EPI == EPI_BIAS_ACT)
                        epi_bias_act(p, accv, row, nt * 256 + c);
                    else
                        epi_gate_residual(p, accv, row, nt * 256 + c);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
    }

    tc_fence_before();
    cluster_sync();  // neither CTA may free TMEM or exit while the other still reads its shared memory / signals its barriers
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2cta(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------- host side
#ifdef TG_DEVELOPER
static int g_gemm_impl = 2;  // 1 = single-CTA 128 x BLOCK_N tiles, 2 = CTA pairs (256 x 256) where the shape allows
#else
static constexpr int g_gemm_impl = 2;  // the shipped configuration (no process-wide mutable state)
#endif

template <int EPI>
static int launch_gemm2(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
    auto kern = gemm2_kernel<EPI>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, g2_smem_bytes(EPI));
        if (e != cudaSuccess) return fail(int(e), "gemm2: cudaFuncSetAttribute(smem=%d): %s", g2_smem_bytes(EPI), cudaGetErrorString(e));
        attr_set = true;
    }
    const int tiles = p.m_tiles * p.n_tiles;
    const int clusters = tiles < sm_count() / 2 ? tiles : sm_count() / 2;
    kern<<<2 * clusters, 256, g2_smem_bytes(EPI), stream>>>(ta, tb, p);
    return check_launch("gemm2");
}

template <int BLOCK_N, int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
    using Cfg = GemmCfg<BLOCK_N>;
    auto kern = gemm_kernel<BLOCK_N, EPI>;
    static bool attr_set = false;  // per instantiation
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return fail(int(e), "gemm: cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
        attr_set = true;
    }
    const int tiles = p.m_tiles * p.n_tiles;
    const int grid = tiles < sm_count() ? tiles : sm_count();
    kern<<<grid, 256, Cfg::SMEM_BYTES, stream>>>(ta, tb, p);
    return check_launch("gemm");
}

static int check_common(const void* A, int64_t lda, const void* W, int M, int N, int K) {
    if (A == nullptr || W == nullptr) return fail(-1, "gemm: null operand");
    if (M <= 0 || N <= 0 || K <= 0) return fail(-2, "gemm: non-positive shape M=%d N=%d K=%d", M, N, K);
    if (K % BLOCK_K != 0) return fail(-3, "gemm: K=%d must be a multiple of %d", K, BLOCK_K);
    if (N % 64 != 0) return fail(-4, "gemm: N=%d must be a multiple of 64", N);
    if (lda % 8 != 0 || lda < K) return fail(-5, "gemm: lda=%lld must be >= K and a multiple of 8", (long long)lda);
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15))
        return fail(-6, "gemm: operands must be 16-byte aligned");
    return 0;
}

template <int EPI>
static int dispatch(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* W, GemmParams& p, cudaStream_t stream) {
    if (g_gemm_impl == 2 && p.N % 256 == 0 && p.M >= 256) {
        p.m_tiles = (p.M + 255) / 256;
        p.n_tiles = p.N / 256;
        CUtensorMap ta, tb;
        int rc = make_tmap_2d(&ta, A, uint64_t(p.K), uint64_t(p.M), uint64_t(lda) * 2, BLOCK_K, 128);
        if (rc) return rc;
        rc = make_tmap_2d(&tb, W, uint64_t(p.K), uint64_t(p.N), uint64_t(p.K) * 2, BLOCK_K, 128);
        if (rc) return rc;
        return launch_gemm2<EPI>(ta, tb, p, stream);
    }
    const int bn = epi_is_qkv(EPI) ? 256 : (p.N % 256 == 0 ? 256 : (p.N % 128 == 0 ? 128 : 64));
    p.m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
    p.n_tiles = p.N / bn;
    CUtensorMap ta, tb;
    int rc = make_tmap_2d(&ta, A, uint64_t(p.K), uint64_t(p.M), uint64_t(lda) * 2, BLOCK_K, BLOCK_M);
    if (rc) return rc;
    rc = make_tmap_2d(&tb, W, uint64_t(p.K), uint64_t(p.N), uint64_t(p.K) * 2, BLOCK_K, uint32_t(bn));
    if (rc) return rc;
    if (bn == 256) return launch_gemm<256, EPI>(ta, tb, p, stream);
    if constexpr (!epi_is_qkv(EPI)) {
        if (bn == 128) return launch_gemm<128, EPI>(ta, tb, p, stream);
        return launch_gemm<64, EPI>(ta, tb, p, stream);
    }
    return fail(-7, "gemm: unsupported tile");
}

}  // namespace tg

using namespace tg;

extern "C" int tg_gemm_bias_act(const tg_bf16* A, int64_t lda, const tg_bf16* W, const tg_bf16* bias, tg_bf16* out,
                                int64_t ldo, int M, int N, int K, int act, void* stream) {
    int rc = check_common(A, lda, W, M, N, K);
    if (rc) return rc;
    if (out == nullptr || ldo % 8 != 0 || ldo < N || (reinterpret_cast<uintptr_t>(out) & 15))
        return fail(-8, "gemm_bias_act: bad output (ldo=%lld)", (long long)ldo);
    if (act < TG_ACT_NONE || act > TG_ACT_SILU) return fail(-9, "gemm_bias_act: unknown activation %d", act);
    GemmParams p{};
    p.M = M; p.N = N; p.K = K;
    p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    p.ldo = ldo;
    p.act = act;
    return dispatch<EPI_BIAS_ACT>(reinterpret_cast<const __nv_bfloat16*>(A), lda,
                                  reinterpret_cast<const __nv_bfloat16*>(W), p, static_cast<cudaStream_t>(stream));
}

static int check_rowmap(const tg_rowmap* map) {
    if (map == nullptr) return fail(-10, "rowmap is null");
    if (map->n_text < 0 || map->n_video < 0 || map->n_vip < 0 ||
        map->rows_per_batch != map->n_text + map->n_video + map->n_vip || map->rows_per_batch <= 0)
        return fail(-11, "rowmap: rows_per_batch must equal n_text+n_video+n_vip");
    if (map->n_video > 0 && (map->hw <= 0 || map->frames <= 0 || map->n_video != map->hw * map->frames))
        return fail(-12, "rowmap: n_video must equal hw*frames");
    if (map->frames <= 0) return fail(-13, "rowmap: frames must be positive");
    if (!rowmap_shard_ok(map)) return fail(-13, "rowmap: shard [row0, row0+rows_local) outside the batch");
    return 0;
}

extern "C" int tg_gemm_gate_residual(const tg_bf16* A, int64_t lda, const tg_bf16* W, const tg_bf16* bias, tg_bf16* X,
                                     int64_t ldx, int B, int N, int K, const tg_rowmap* map, const tg_modvec* gate,
                                     void* stream) {
    int rc = check_rowmap(map);
    if (rc) return rc;
    if (B <= 0) return fail(-2, "gemm_gate_residual: B=%d", B);
    const int M = B * normalised_rowmap(map).rows_local;
    rc = check_common(A, lda, W, M, N, K);
    if (rc) return rc;
    if (X == nullptr || ldx % 8 != 0 || ldx < N || (reinterpret_cast<uintptr_t>(X) & 15))
        return fail(-8, "gemm_gate_residual: bad X (ldx=%lld)", (long long)ldx);
    if (gate == nullptr) return fail(-14, "gemm_gate_residual: gate is null");
    if ((map->n_text > 0 && !gate->text) || (map->n_video > 0 && !gate->video) || (map->n_vip > 0 && !gate->vip))
        return fail(-15, "gemm_gate_residual: a populated segment has no gate vector");
    GemmParams p{};
    p.M = M; p.N = N; p.K = K;
    p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
    p.out = reinterpret_cast<__nv_bfloat16*>(X);
    p.ldo = ldx;
    p.map = normalised_rowmap(map);
    p.gate = *gate;
    return dispatch<EPI_GATE_RESIDUAL>(reinterpret_cast<const __nv_bfloat16*>(A), lda,
                                       reinterpret_cast<const __nv_bfloat16*>(W), p,
                                       static_cast<cudaStream_t>(stream));
}

static int qkv_rope_gemm_impl(const tg_bf16* A, int64_t lda, const tg_bf16* W, const tg_bf16* bias, int B, int H,
                              int K, const tg_rowmap* map, const tg_qkv_proj* proj, int nproj, float ln_eps,
                              const tg_qkv_scatter* scatter, void* stream) {
    int rc = check_rowmap(map);
    if (rc) return rc;
    if (B <= 0 || H <= 0) return fail(-2, "qkv_rope_gemm: B=%d H=%d", B, H);
    if (nproj < 1 || nproj > 6 || proj == nullptr) return fail(-16, "qkv_rope_gemm: nproj=%d (1..6)", nproj);
    if ((H * 64) % 256 != 0) return fail(-17, "qkv_rope_gemm: H*64 must be a multiple of 256");
    const int M = B * normalised_rowmap(map).rows_local;
    const int N = nproj * H * 64;
    rc = check_common(A, lda, W, M, N, K);
    if (rc) return rc;
    GemmParams p{};
    p.M = M; p.N = N; p.K = K;
    p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
    p.map = normalised_rowmap(map);
    p.heads = H;
    p.ln_eps = ln_eps;
    if (scatter != nullptr) {
        if (scatter->world < 1 || scatter->world > TG_MAX_PEERS || H % scatter->world != 0)
            return fail(-20, "qkv_rope_gemm_sp: world=%d must be in 1..%d and divide H=%d", scatter->world, TG_MAX_PEERS, H);
        p.sp_world = scatter->world;
        p.sp_hpr = H / scatter->world;
        for (int i = 0; i < nproj; ++i)
            for (int q = 0; q < scatter->world; ++q) {
                if (scatter->peer[i][q] == nullptr || (reinterpret_cast<uintptr_t>(scatter->peer[i][q]) & 15))
                    return fail(-21, "qkv_rope_gemm_sp: projection %d has no (16-byte aligned) buffer on rank %d", i, q);
                p.sp_peer[i][q] = reinterpret_cast<__nv_bfloat16*>(scatter->peer[i][q]);
            }
    }
    for (int i = 0; i < nproj; ++i) {
        if ((scatter == nullptr && proj[i].out == nullptr) || proj[i].out_rows <= 0 || proj[i].out_rows > map->rows_per_batch)
            return fail(-18, "qkv_rope_gemm: projection %d has a bad output", i);
        if ((proj[i].ln_w == nullptr) != (proj[i].ln_b == nullptr) ||
            (proj[i].cos_video == nullptr) != (proj[i].sin_video == nullptr) ||
            (proj[i].cos_vip == nullptr) != (proj[i].sin_vip == nullptr))
            return fail(-19, "qkv_rope_gemm: projection %d has half of a (weight,bias) or (cos,sin) pair", i);
        p.proj[i].out = reinterpret_cast<__nv_bfloat16*>(scatter != nullptr ? scatter->peer[i][0] : proj[i].out);  // world 1
        p.proj[i].out_rows = proj[i].out_rows;
        p.proj[i].ln_w = reinterpret_cast<const __nv_bfloat16*>(proj[i].ln_w);
        p.proj[i].ln_b = reinterpret_cast<const __nv_bfloat16*>(proj[i].ln_b);
        p.proj[i].cos_video = proj[i].cos_video;
        p.proj[i].sin_video = proj[i].sin_video;
        p.proj[i].cos_vip = proj[i].cos_vip;
        p.proj[i].sin_vip = proj[i].sin_vip;
    }
    if (p.sp_world > 1)
        return dispatch<EPI_QKV_SP>(reinterpret_cast<const __nv_bfloat16*>(A), lda, reinterpret_cast<const __nv_bfloat16*>(W),
                                    p, static_cast<cudaStream_t>(stream));
    return dispatch<EPI_QKV>(reinterpret_cast<const __nv_bfloat16*>(A), lda, reinterpret_cast<const __nv_bfloat16*>(W),
                             p, static_cast<cudaStream_t>(stream));
}

extern "C" int tg_qkv_rope_gemm(const tg_bf16* A, int64_t lda, const tg_bf16* W, const tg_bf16* bias, int B, int H,
                                int K, const tg_rowmap* map, const tg_qkv_proj* proj, int nproj, float ln_eps,
                                void* stream) {
    return qkv_rope_gemm_impl(A, lda, W, bias, B, H, K, map, proj, nproj, ln_eps, nullptr, stream);
}

extern "C" int tg_qkv_rope_gemm_sp(const tg_bf16* A, int64_t lda, const tg_bf16* W, const tg_bf16* bias, int B, int H,
                                   int K, const tg_rowmap* map, const tg_qkv_proj* proj, int nproj, float ln_eps,
                                   const tg_qkv_scatter* scatter, void* stream) {
    if (scatter == nullptr) return fail(-20, "qkv_rope_gemm_sp: scatter is null");
    return qkv_rope_gemm_impl(A, lda, W, bias, B, H, K, map, proj, nproj, ln_eps, scatter, stream);
}

#ifdef TG_DEVELOPER
extern "C" int tg_set_gemm_impl(int impl) {  // developer hook (1 = single-CTA tiles, 2 = CTA pairs); not in the public header
    if (impl != 1 && impl != 2) return tg::fail(-1, "gemm impl must be 1 or 2");
    tg::g_gemm_impl = impl;
    return 0;
}
#endif
