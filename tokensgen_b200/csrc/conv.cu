// tcgen05 implicit-GEMM convolution for the 3D causal VAE: CogVideoXCausalConv3d (3x3x3), the per-frame
// 3x3 convolutions of the up/down-samplers (stride 1 or 2) and 1x1x1 projections, on channels-last activations.
//
//   Y[t, h, w, n] = bias[n] + sum_{kt,kh,kw,c} X[t + kt, h*s + kh - ph, w*s + kw - pw, c] * W[n, (kt,kh,kw,c)]  (+ R[t,h,w,n])
//
// X already carries the kt-1 causal frames in front (conv cache or first-frame copies, written by the caller), so the time
// axis needs no padding; the zero padding in H and W is the TMA's out-of-bounds fill — no im2col buffer, no F.pad, no cat.
//
// One persistent CTA per SM, 256 threads, same warp roles as gemm.cu:
//   warp 0  TMA producer : per k-block (one tap x 64 input channels) ONE 4-D box {64 c, 32 w, 8 h, 1 t} of X (32 KB, the
//                          256 output pixels of the tile shifted by the tap) + a {64 k, BLOCK_N} box of W
//   warp 1  MMA issuer   : the 256-pixel tile is two 128-row UMMA operands (h rows 0-3 / 4-7): 2 x 4 tcgen05.mma per stage
//   warp 2  TMEM alloc   : 2 accumulator buffers x 2 halves x BLOCK_N columns
//   warps 4..7 epilogue  : thread = pixel; bias, optional residual, bf16 store channels-last (or fp32/bf16 planes)
// Tile = 256 pixels x BLOCK_N channels: 48 KB of operands per 256x128x64 MACs, the same operand traffic per flop as the
// 128x256 GEMM tile that runs at 96 % of the cuBLAS peak.
#include <cuda_bf16.h>

#include "common.h"
#include "ptx.cuh"

namespace tg {

constexpr int CV_TW = 32, CV_TH = 8;           // output patch of one tile
constexpr int CV_A_BYTES = CV_TW * CV_TH * 64 * 2;  // 32 KB

struct ConvParams {
    int T_out, H_out, W_out, Cout;
    int kt, kh, kw, stride, pad_h, pad_w;
    int c_chunks;             // Cin / 64
    int h_tiles, w_tiles, n_tiles;
    const __nv_bfloat16* bias;
    const __nv_bfloat16* residual;  // channels-last [T_out, H_out, W_out, ld_res] or null
    int64_t ld_res;
    __nv_bfloat16* y;
    int64_t ldy;              // layout 0: channel stride between pixels
    int64_t plane_stride;     // layout 1: elements between channel planes; pixel (t,h,w) at (t*H_out + h)*W_out + w
    int layout;               // 0 channels-last, 1 channel planes
    // GroupNorm statistics of the OUTPUT, accumulated in the epilogue (replaces the tg_vae_group_stats pass over the tensor the
    // next GroupNorm / SpatialNorm reads): stats[g] += sum, stats[groups + g] += sum of squares of the bf16-ROUNDED outputs of
    // group g's Cout / groups channels.  Null = off.  Channels-last layout with Cout % 32 == 0 and 32 % (Cout / groups) == 0 or
    // (Cout / groups) % 32 == 0 only (checked by the host).
    double* stats;
    int stat_groups;
};

constexpr int CV_MAX_GROUPS = 64;

// Per-chunk GroupNorm partials: v[32] are this pixel's bf16-rounded outputs of channels [n0, n0 + 32).  The warp's 32 pixels are
// summed with a transpose-reduce butterfly (2 NG values in log2(2 NG) halving steps + the remaining full steps: 2 NG shuffles in
// all instead of 5 per value), which leaves value i on the lanes whose upper bits spell i; those lanes add it to THIS WARP's
// row of the CTA's accumulators (sh[warp][g] sums, sh[warp][G + g] squares; plain read-modify-write, one lane per slot, program
// order — no atomics, so a CTA's partial sums do not depend on how its four epilogue warps interleave and the statistics are
// reproducible run to run; the rows are folded in fp64 and flushed with fp64 atomics once per CTA).
template <int CG>   // channels per group inside the chunk: 4, 8, 16 or 32 (= the whole chunk belongs to one group)
__device__ __forceinline__ void conv_chunk_stats(const float (&v)[32], int n0, int cg_total, int groups, float* sh) {
    constexpr int NG = 32 / CG;
    constexpr int NV = 2 * NG;       // 16, 8, 4 or 2
    float a[NV];
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        float s = 0.f, q = 0.f;
#pragma unroll
        for (int i = 0; i < CG; ++i) {
            s += v[g * CG + i];
            q = fmaf(v[g * CG + i], v[g * CG + i], q);
        }
        a[g] = s;
        a[NG + g] = q;
    }
    const int lane = threadIdx.x & 31;
    int m = 16;
#pragma unroll
    for (int n = NV; n > 1; n >>= 1, m >>= 1) {
        const bool hi = (lane & m) != 0;
#pragma unroll
        for (int i = 0; i < n / 2; ++i) {
            const float send = hi ? a[i] : a[i + n / 2];
            const float keep = hi ? a[i + n / 2] : a[i];
            a[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
        }
    }
#pragma unroll
    for (; m > 0; m >>= 1) a[0] += __shfl_xor_sync(0xffffffffu, a[0], m);
    constexpr int SHIFT = (NV == 16) ? 1 : (NV == 8) ? 2 : (NV == 4) ? 3 : 4;   // 5 - log2(NV)
    if ((lane & ((1 << SHIFT) - 1)) == 0) {
        const int idx = lane >> SHIFT;                 // which of the NV values this lane holds
        const int g0 = n0 / cg_total;                  // first group of this chunk (CG == 32: the group the chunk lies in)
        const int slot = idx < NG ? g0 + idx : groups + g0 + (idx - NG);
        float* mine = sh + ((threadIdx.x >> 5) & 3) * (2 * CV_MAX_GROUPS) + slot;
        *mine += a[0];
    }
    __syncwarp();
}

template <int BLOCK_N>
struct ConvCfg {
    static constexpr int B_BYTES = BLOCK_N * 64 * 2;
    static constexpr int STAGE_BYTES = CV_A_BYTES + B_BYTES;
    static constexpr int STAGES = (BLOCK_N == 128) ? 4 : 5;
    static constexpr int TMEM_COLS = 4 * BLOCK_N;  // 512 / 256
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(m), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
}

__device__ __forceinline__ void conv_tile_coords(const ConvParams& p, int tile, int& nt, int& wt, int& ht, int& t) {
    nt = tile % p.n_tiles;  // channel tiles of one pixel tile run side by side: they share the activation boxes in L2
    int r = tile / p.n_tiles;
    wt = r % p.w_tiles;
    r /= p.w_tiles;
    ht = r % p.h_tiles;
    t = r / p.h_tiles;
}

// Epilogue of one accumulator row (one output pixel): BLOCK_N columns starting at output channel nt * BLOCK_N.
// Called by whole warps (all 32 lanes, `ok` may differ per lane): the statistics path shuffles.
template <int BLOCK_N>
__device__ __forceinline__ void conv_epilogue_row(const ConvParams& p, uint32_t t_row, int nt, bool ok, int64_t pix,
                                                  float* sh_stats) {
#pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += 32) {
            uint32_t r[32];
            tmem_ld32(t_row + c, r);
            tmem_wait_ld();
            const int n0 = nt * BLOCK_N + c;
            if (n0 >= p.Cout) continue;                       // warp-uniform
            const bool fast = p.layout == 0 && n0 + 32 <= p.Cout;
            if (!ok && !(fast && p.stats != nullptr)) continue;
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
            if (fast) {
                if (ok) {
                if (p.bias != nullptr) {
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        const uint4 bv = __ldg(reinterpret_cast<const uint4*>(p.bias + n0 + i));
                        v[i] += bf16_lo(bv.x); v[i + 1] += bf16_hi(bv.x); v[i + 2] += bf16_lo(bv.y); v[i + 3] += bf16_hi(bv.y);
                        v[i + 4] += bf16_lo(bv.z); v[i + 5] += bf16_hi(bv.z); v[i + 6] += bf16_lo(bv.w); v[i + 7] += bf16_hi(bv.w);
                    }
                }
                if (p.residual != nullptr) {
                    const uint4* rp = reinterpret_cast<const uint4*>(p.residual + pix * p.ld_res + n0);
                    uint4 rv[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) rv[i] = rp[i];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        v[i * 8] += bf16_lo(rv[i].x); v[i * 8 + 1] += bf16_hi(rv[i].x);
                        v[i * 8 + 2] += bf16_lo(rv[i].y); v[i * 8 + 3] += bf16_hi(rv[i].y);
                        v[i * 8 + 4] += bf16_lo(rv[i].z); v[i * 8 + 5] += bf16_hi(rv[i].z);
                        v[i * 8 + 6] += bf16_lo(rv[i].w); v[i * 8 + 7] += bf16_hi(rv[i].w);
                    }
                }
                uint4* yp = reinterpret_cast<uint4*>(p.y + pix * p.ldy + n0);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint4 o;
                    o.x = pack_bf16x2(v[i * 8], v[i * 8 + 1]); o.y = pack_bf16x2(v[i * 8 + 2], v[i * 8 + 3]);
                    o.z = pack_bf16x2(v[i * 8 + 4], v[i * 8 + 5]); o.w = pack_bf16x2(v[i * 8 + 6], v[i * 8 + 7]);
                    yp[i] = o;
                    if (p.stats != nullptr) {   // the statistics are those of the STORED (bf16) tensor, like the separate pass
                        v[i * 8] = bf16_lo(o.x); v[i * 8 + 1] = bf16_hi(o.x); v[i * 8 + 2] = bf16_lo(o.y); v[i * 8 + 3] = bf16_hi(o.y);
                        v[i * 8 + 4] = bf16_lo(o.z); v[i * 8 + 5] = bf16_hi(o.z); v[i * 8 + 6] = bf16_lo(o.w); v[i * 8 + 7] = bf16_hi(o.w);
                    }
                }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0.f;   // pixel outside the image: contributes nothing
                }
                if (p.stats != nullptr) {
                    const int cg = p.Cout / p.stat_groups;
                    switch (cg) {
                        case 4: conv_chunk_stats<4>(v, n0, cg, p.stat_groups, sh_stats); break;
                        case 8: conv_chunk_stats<8>(v, n0, cg, p.stat_groups, sh_stats); break;
                        case 16: conv_chunk_stats<16>(v, n0, cg, p.stat_groups, sh_stats); break;
                        default: conv_chunk_stats<32>(v, n0, cg, p.stat_groups, sh_stats); break;   // cg % 32 == 0
                    }
                }
            } else {
                // narrow outputs (conv_out: 3 image channels / 32 moment channels) and channel-plane layout
                for (int i = 0; i < 32 && n0 + i < p.Cout; ++i) {
                    float o = v[i] + (p.bias != nullptr ? __bfloat162float(p.bias[n0 + i]) : 0.f);
                    if (p.residual != nullptr) o += __bfloat162float(p.residual[pix * p.ld_res + n0 + i]);
                    if (p.layout == 0) p.y[pix * p.ldy + n0 + i] = __float2bfloat16_rn(o);
                    else p.y[int64_t(n0 + i) * p.plane_stride + pix] = __float2bfloat16_rn(o);
                }
            }
        }
}

// The CTA's statistics accumulators: zeroed before the tile loop, flushed to p.stats (fp64 atomics) after it.  Called by all
// 128 epilogue threads (warps 4..7); named barrier 1.
__device__ __forceinline__ void conv_stats_begin(const ConvParams& p, float* sh) {
    if (p.stats == nullptr) return;
    for (int i = threadIdx.x - 128; i < 4 * 2 * CV_MAX_GROUPS; i += 128) sh[i] = 0.f;
    named_bar_sync(1, 128);
}
__device__ __forceinline__ void conv_stats_end(const ConvParams& p, float* sh) {
    if (p.stats == nullptr) return;
    named_bar_sync(1, 128);
    for (int i = threadIdx.x - 128; i < 2 * p.stat_groups; i += 128) {
        const double v = double(sh[i]) + double(sh[2 * CV_MAX_GROUPS + i]) + double(sh[4 * CV_MAX_GROUPS + i]) +
                         double(sh[6 * CV_MAX_GROUPS + i]);
        if (v != 0.0) atomicAdd(&p.stats[i], v);
    }
}

template <int BLOCK_N>
__global__ void __launch_bounds__(256, 1)
conv_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
            const __grid_constant__ ConvParams p) {
    using Cfg = ConvCfg<BLOCK_N>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    __shared__ float sh_stats[4 * 2 * CV_MAX_GROUPS];   // one row per epilogue warp
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
    const int num_tiles = p.T_out * p.h_tiles * p.w_tiles * p.n_tiles;
    const int taps = p.kt * p.kh * p.kw;
    const int k_blocks = taps * p.c_chunks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 4);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int nt, wt, ht, t;
                conv_tile_coords(p, tile, nt, wt, ht, t);
                const int w_in0 = wt * CV_TW * p.stride - p.pad_w;
                const int h_in0 = ht * CV_TH * p.stride - p.pad_h;
                int kb = 0;
                for (int it = 0; it < p.kt; ++it)
                    for (int ih = 0; ih < p.kh; ++ih)
                        for (int iw = 0; iw < p.kw; ++iw)
                            for (int cc = 0; cc < p.c_chunks; ++cc, ++kb) {
                                mbar_wait(empty_bar(stage), phase ^ 1u, 0x401);
                                const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                                mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
                                tma_load_4d(sa, &tmap_x, full_bar(stage), cc * 64, w_in0 + iw, h_in0 + ih, t + it);
                                tma_load_2d(sa + CV_A_BYTES, &tmap_w, full_bar(stage), kb * 64, nt * BLOCK_N);
                                if (++stage == STAGES) {
                                    stage = 0;
                                    phase ^= 1u;
                                }
                            }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(128, BLOCK_N, false, false);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u, 0x402);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + uint32_t(acc * 2 * BLOCK_N);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(full_bar(stage), phase, 0x403);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t sb = sa + CV_A_BYTES;
#pragma unroll
                    for (int half = 0; half < 2; ++half)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t da = make_smem_desc_sw128(sa + half * (CV_A_BYTES / 2) + k * 32, 16, 1024);
                            const uint64_t db = make_smem_desc_sw128(sb + k * 32, 16, 1024);
                            umma_ss(d_tmem + uint32_t(half * BLOCK_N), da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                    umma_commit(empty_bar(stage));
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit(tfull_bar(acc));
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
        }
    } else if (warp >= 4) {
        const int ew = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        conv_stats_begin(p, sh_stats);
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int nt, wt, ht, t;
            conv_tile_coords(p, tile, nt, wt, ht, t);
            mbar_wait(tfull_bar(acc), acc_phase, 0x404);
            tc_fence_after();
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int R = half * 128 + ew * 32 + lane;  // row of the 256-pixel tile = th * 32 + tw
                const int h = ht * CV_TH + (R >> 5), w = wt * CV_TW + (R & 31);
                const bool ok = h < p.H_out && w < p.W_out;
                const int64_t pix = (int64_t(t) * p.H_out + h) * p.W_out + w;
                const uint32_t t_row = tmem_base + (uint32_t(ew * 32) << 16) + uint32_t(acc * 2 * BLOCK_N + half * BLOCK_N);
                conv_epilogue_row<BLOCK_N>(p, t_row, nt, ok, pix, sh_stats);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
        conv_stats_end(p, sh_stats);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// =====================================================================================================================
// CTA-pair convolution (cta_group::2).  The single-CTA kernel above is bound by operand delivery, not by the tensor pipe
// (ncu: 29.3 GB through the L2 -> SM crossbar for one 8 x 480 x 720 x 128 -> 128 layer = 13.1 TB/s, tensor pipe 58 %).
// Here the two CTAs of a cluster take two DIFFERENT 256-pixel tiles and the SAME BLOCK_N output channels: each loads its
// own activation boxes and HALF of the weight tile, and the leader's MMA thread issues M = 256 (128 pixels of each CTA)
// x N = BLOCK_N instructions, so the weight operand crosses L2 -> SM once per pair.  Operand bytes per 256 x 128 x 64 MACs:
// 48 KB (single CTA) -> 40 KB (BLOCK_N = 128) -> 24 KB (BLOCK_N = 256, one accumulator buffer: 2 x 256 TMEM columns per
// pixel half).  Barrier protocol as in gemm2_kernel.
template <int BLOCK_N>
struct Conv2Cfg {
    static constexpr int HALF_N = BLOCK_N / 2;
    static constexpr int B_BYTES = HALF_N * 64 * 2;
    static constexpr int STAGE_BYTES = CV_A_BYTES + B_BYTES;  // 40 / 48 KB
    static constexpr int STAGES = (BLOCK_N == 128) ? 5 : 4;   // 200 / 192 KB
    static constexpr int ACC_BUFS = (BLOCK_N == 128) ? 2 : 1;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;
};

template <int BLOCK_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
conv2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
             const __grid_constant__ ConvParams p) {
    using Cfg = Conv2Cfg<BLOCK_N>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    __shared__ float sh_stats[4 * 2 * CV_MAX_GROUPS];   // one row per epilogue warp
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
    const int rank = int(cluster_ctarank());
    const bool leader = rank == 0;
    const int n_clusters = gridDim.x >> 1;
    const int cluster_id = blockIdx.x >> 1;
    const int pixel_tiles = p.T_out * p.h_tiles * p.w_tiles;
    const int num_tiles = ((pixel_tiles + 1) >> 1) * p.n_tiles;  // (pair of pixel tiles) x channel tile
    const int taps = p.kt * p.kh * p.kw;
    const int k_blocks = taps * p.c_chunks;
    // this CTA's pixel tile of pair tile `tile`; t == T_out marks the padding tile of an odd count (loads zero-fill, stores are masked)
    auto my_coords = [&](int tile, int& nt, int& wt, int& ht, int& t) {
        nt = tile % p.n_tiles;
        const int q = 2 * (tile / p.n_tiles) + rank;
        wt = q % p.w_tiles;
        const int r = q / p.w_tiles;
        ht = r % p.h_tiles;
        t = r / p.h_tiles;
    };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 8);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_2cta(tmem_slot, 512);
    tc_fence_before();
    cluster_sync();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
                int nt, wt, ht, t;
                my_coords(tile, nt, wt, ht, t);
                const int w_in0 = wt * CV_TW * p.stride - p.pad_w;
                const int h_in0 = ht * CV_TH * p.stride - p.pad_h;
                int kb = 0;
                for (int it = 0; it < p.kt; ++it)
                    for (int ih = 0; ih < p.kh; ++ih)
                        for (int iw = 0; iw < p.kw; ++iw)
                            for (int cc = 0; cc < p.c_chunks; ++cc, ++kb) {
                                mbar_wait(empty_bar(stage), phase ^ 1u, 0x481);
                                const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                                const uint32_t bar = mapa_shared(full_bar(stage), 0);
                                if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);
                                tma_load_4d_2cta(sa, &tmap_x, bar, cc * 64, w_in0 + iw, h_in0 + ih, t + it);
                                tma_load_2d_2cta(sa + CV_A_BYTES, &tmap_w, bar, kb * 64, nt * BLOCK_N + rank * Cfg::HALF_N);
                                if (++stage == STAGES) {
                                    stage = 0;
                                    phase ^= 1u;
                                }
                            }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            constexpr uint32_t idesc = make_idesc_bf16(256, BLOCK_N, false, false);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u, 0x482);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + uint32_t(acc * 2 * BLOCK_N);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(full_bar(stage), phase, 0x483);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t sb = sa + CV_A_BYTES;
#pragma unroll
                    for (int half = 0; half < 2; ++half)
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t da = make_smem_desc_sw128(sa + half * (CV_A_BYTES / 2) + k * 32, 16, 1024);
                            const uint64_t db = make_smem_desc_sw128(sb + k * 32, 16, 1024);
                            umma_ss_2cta(d_tmem + uint32_t(half * BLOCK_N), da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                    umma_commit_2cta(empty_bar(stage), 0b11);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit_2cta(tfull_bar(acc), 0b11);
                if (++acc == Cfg::ACC_BUFS) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
        }
    } else if (warp >= 4) {
        const int ew = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        conv_stats_begin(p, sh_stats);
        for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
            int nt, wt, ht, t;
            my_coords(tile, nt, wt, ht, t);
            mbar_wait(tfull_bar(acc), acc_phase, 0x484);
            tc_fence_after();
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int R = half * 128 + ew * 32 + lane;
                const int h = ht * CV_TH + (R >> 5), w = wt * CV_TW + (R & 31);
                const bool ok = t < p.T_out && h < p.H_out && w < p.W_out;
                const int64_t pix = (int64_t(t) * p.H_out + h) * p.W_out + w;
                const uint32_t t_row = tmem_base + (uint32_t(ew * 32) << 16) + uint32_t(acc * 2 * BLOCK_N + half * BLOCK_N);
                conv_epilogue_row<BLOCK_N>(p, t_row, nt, ok, pix, sh_stats);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));
            if (++acc == Cfg::ACC_BUFS) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
        conv_stats_end(p, sh_stats);
    }

    tc_fence_before();
    cluster_sync();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2cta(tmem_base, 512);
    }
}

// =====================================================================================================================
// Tap-reuse CTA-pair convolution (stride 1, kw = 3).  conv2 still streams every activation box once per TAP: 27 boxes of
// 32 KB per 256-pixel tile and 64 input channels.  The three kw taps of one (kt, kh) read the same pixels shifted by one
// pixel along w, so here the tile is 128 w x 2 h (each h row = one 128-row UMMA operand, contiguous in w) and ONE halo'd
// box {64 c, 130 w, 2 h} per (kt, kh, channel chunk) serves all three kw taps: the A descriptor of tap kw simply starts
// kw rows (kw x 128 B) further into the box.  That start is not at a 1024-byte boundary of the 128-byte swizzle pattern;
// measured on B200 (tests/test_vae_gpu.py::test_conv_tap_reuse_kernel): the tensor core applies the swizzle XOR to the
// ABSOLUTE shared-memory address bits, exactly as the TMA did when it wrote the box, so a plain descriptor with the shifted
// start address is correct and the descriptor's base-offset field must stay 0 (setting it to the swizzle phase
// double-corrects and gives wrong sums).  Activation traffic drops 3x
// (9 boxes of 33 KB instead of 27 of 32 KB); with the weight tile split across the CTA pair a 256-pixel x 128-channel x
// 64-deep unit moves ~19 KB instead of 40 KB (conv2<128>) / 48 KB (single CTA).
constexpr int C3_TW = 128, C3_TH = 2;
constexpr int C3_ROW_BYTES = (C3_TW + 2) * 128;                       // one h row of the halo'd box: 130 pixels x 64 c
constexpr int C3_A_BYTES = C3_TH * C3_ROW_BYTES;                       // 33 280 B delivered by the TMA
constexpr int C3_A_SLOT = (C3_A_BYTES + 1023) / 1024 * 1024;           // 33 792 B
// BLOCK_N output channels per tile (128; 32 for the narrow conv_out, whose MMAs at N = 128 would cost more than its boxes):
// each CTA of the pair loads BLOCK_N / 2 weight rows per tap
template <int BLOCK_N>
struct Conv3Cfg {
    static constexpr int HALF_N = BLOCK_N / 2;
    static constexpr int B_BYTES = HALF_N * 64 * 2;                      // 8 KB (2 KB) per tap
    static constexpr int STAGE_BYTES = C3_A_SLOT + 3 * B_BYTES;         // 58 368 B (39 936 B)
    static constexpr int STAGE_TX = C3_A_BYTES + 3 * B_BYTES;           // bytes the TMA actually delivers per CTA and stage
    static constexpr int STAGES = BLOCK_N == 128 ? 3 : 5;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;
};

template <int BLOCK_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1)
conv3_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
             const __grid_constant__ ConvParams p) {
    using Cfg = Conv3Cfg<BLOCK_N>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    __shared__ float sh_stats[4 * 2 * CV_MAX_GROUPS];   // one row per epilogue warp
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
    const int rank = int(cluster_ctarank());
    const bool leader = rank == 0;
    const int n_clusters = gridDim.x >> 1;
    const int cluster_id = blockIdx.x >> 1;
    const int pixel_tiles = p.T_out * p.h_tiles * p.w_tiles;       // h_tiles / w_tiles in units of the 128 x 2 tile
    const int num_tiles = ((pixel_tiles + 1) >> 1) * p.n_tiles;
    const int k_blocks = p.kt * p.kh * p.c_chunks;                 // one k-block = (kt, kh, channel chunk): three kw taps
    auto my_coords = [&](int tile, int& nt, int& wt, int& ht, int& t) {
        nt = tile % p.n_tiles;
        const int q = 2 * (tile / p.n_tiles) + rank;
        wt = q % p.w_tiles;
        const int r = q / p.w_tiles;
        ht = r % p.h_tiles;
        t = r / p.h_tiles;
    };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_w);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), 8);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc_2cta(tmem_slot, 512);
    tc_fence_before();
    cluster_sync();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
                int nt, wt, ht, t;
                my_coords(tile, nt, wt, ht, t);
                const int w_in0 = wt * C3_TW - p.pad_w;
                const int h_in0 = ht * C3_TH - p.pad_h;
                for (int it = 0; it < p.kt; ++it)
                    for (int ih = 0; ih < p.kh; ++ih)
                        for (int cc = 0; cc < p.c_chunks; ++cc) {
                            mbar_wait(empty_bar(stage), phase ^ 1u, 0x4a1);
                            const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                            const uint32_t bar = mapa_shared(full_bar(stage), 0);
                            if (leader) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::STAGE_TX);
                            tma_load_4d_2cta(sa, &tmap_x, bar, cc * 64, w_in0, h_in0 + ih, t + it);
#pragma unroll
                            for (int iw = 0; iw < 3; ++iw) {
                                const int kb = ((it * p.kh + ih) * 3 + iw) * p.c_chunks + cc;
                                tma_load_2d_2cta(sa + C3_A_SLOT + iw * Cfg::B_BYTES, &tmap_w, bar, kb * 64,
                                                 nt * BLOCK_N + rank * Cfg::HALF_N);
                            }
                            if (++stage == STAGES) {
                                stage = 0;
                                phase ^= 1u;
                            }
                        }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && leader) {
            constexpr uint32_t idesc = make_idesc_bf16(256, BLOCK_N, false, false);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u, 0x4a2);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + uint32_t(acc * 2 * BLOCK_N);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(full_bar(stage), phase, 0x4a3);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t sb = sa + C3_A_SLOT;
#pragma unroll
                    for (int row = 0; row < C3_TH; ++row)
#pragma unroll
                        for (int iw = 0; iw < 3; ++iw)
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t da = make_smem_desc_sw128(sa + row * C3_ROW_BYTES + iw * 128 + k * 32, 16, 1024);
                                const uint64_t db = make_smem_desc_sw128(sb + iw * Cfg::B_BYTES + k * 32, 16, 1024);
                                umma_ss_2cta(d_tmem + uint32_t(row * BLOCK_N), da, db, idesc, (kb | iw | k) != 0 ? 1u : 0u);
                            }
                    umma_commit_2cta(empty_bar(stage), 0b11);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit_2cta(tfull_bar(acc), 0b11);
                if (++acc == 2) {
                    acc = 0;
                    acc_phase ^= 1u;
                }
            }
        }
    } else if (warp >= 4) {
        const int ew = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        conv_stats_begin(p, sh_stats);
        for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
            int nt, wt, ht, t;
            my_coords(tile, nt, wt, ht, t);
            mbar_wait(tfull_bar(acc), acc_phase, 0x4a4);
            tc_fence_after();
#pragma unroll 1
            for (int row = 0; row < C3_TH; ++row) {
                const int h = ht * C3_TH + row, w = wt * C3_TW + ew * 32 + lane;
                const bool ok = t < p.T_out && h < p.H_out && w < p.W_out;
                const int64_t pix = (int64_t(t) * p.H_out + h) * p.W_out + w;
                const uint32_t t_row = tmem_base + (uint32_t(ew * 32) << 16) + uint32_t(acc * 2 * BLOCK_N + row * BLOCK_N);
                conv_epilogue_row<BLOCK_N>(p, t_row, nt, ok, pix, sh_stats);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));
            if (++acc == 2) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
        conv_stats_end(p, sh_stats);
    }

    tc_fence_before();
    cluster_sync();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2cta(tmem_base, 512);
    }
}

template <int BLOCK_N>
static int launch_conv3(const CUtensorMap& tx, const CUtensorMap& tw, const ConvParams& p, cudaStream_t stream) {
    using Cfg = Conv3Cfg<BLOCK_N>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv3_kernel<BLOCK_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return fail(int(e), "vae_conv: cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
        attr_set = true;
    }
    const int tiles = ((p.T_out * p.h_tiles * p.w_tiles + 1) / 2) * p.n_tiles;
    const int clusters = tiles < sm_count() / 2 ? tiles : sm_count() / 2;
    conv3_kernel<BLOCK_N><<<2 * clusters, 256, Cfg::SMEM_BYTES, stream>>>(tx, tw, p);
    return check_launch("vae_conv");
}

#ifdef TG_DEVELOPER
static int g_conv_impl = 3;  // 1 = single-CTA tiles, 2 = CTA pairs, 3 = CTA pairs + kw-tap reuse where stride 1 / kw 3 allow
#else
static constexpr int g_conv_impl = 3;  // the shipped configuration (no process-wide mutable state)
#endif

template <int BLOCK_N>
static int launch_conv2(const CUtensorMap& tx, const CUtensorMap& tw, const ConvParams& p, cudaStream_t stream) {
    using Cfg = Conv2Cfg<BLOCK_N>;
    auto kern = conv2_kernel<BLOCK_N>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return fail(int(e), "vae_conv: cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
        attr_set = true;
    }
    const int tiles = ((p.T_out * p.h_tiles * p.w_tiles + 1) / 2) * p.n_tiles;
    const int clusters = tiles < sm_count() / 2 ? tiles : sm_count() / 2;
    kern<<<2 * clusters, 256, Cfg::SMEM_BYTES, stream>>>(tx, tw, p);
    return check_launch("vae_conv");
}

template <int BLOCK_N>
static int launch_conv(const CUtensorMap& tx, const CUtensorMap& tw, const ConvParams& p, cudaStream_t stream) {
    using Cfg = ConvCfg<BLOCK_N>;
    auto kern = conv_kernel<BLOCK_N>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return fail(int(e), "vae_conv: cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
        attr_set = true;
    }
    const int tiles = p.T_out * p.h_tiles * p.w_tiles * p.n_tiles;
    const int grid = tiles < sm_count() ? tiles : sm_count();
    kern<<<grid, 256, Cfg::SMEM_BYTES, stream>>>(tx, tw, p);
    return check_launch("vae_conv");
}

}  // namespace tg

using namespace tg;

extern "C" int tg_vae_conv(const tg_conv_args* a, void* stream) {
    if (a == nullptr || a->x == nullptr || a->w == nullptr || a->y == nullptr) return fail(-1, "vae_conv: null pointer");
    if (a->Cin <= 0 || a->Cin % 64 != 0) return fail(-2, "vae_conv: Cin=%d must be a positive multiple of 64 (pad the channels)", a->Cin);
    if (a->Cout <= 0 || a->Cout_pad < a->Cout || a->Cout_pad % 64 != 0)
        return fail(-3, "vae_conv: Cout=%d Cout_pad=%d (Cout_pad: multiple of 64, >= Cout)", a->Cout, a->Cout_pad);
    if (a->kt < 1 || a->kt > 3 || a->kh < 1 || a->kh > 3 || a->kw < 1 || a->kw > 3) return fail(-4, "vae_conv: kernel extents must be 1..3");
    if (a->stride_hw != 1 && a->stride_hw != 2) return fail(-5, "vae_conv: stride_hw must be 1 or 2");
    if (a->T_out <= 0 || a->H_out <= 0 || a->W_out <= 0 || a->T_in < a->T_out + a->kt - 1)
        return fail(-6, "vae_conv: T_in=%d must hold T_out=%d + kt-1 causal frames", a->T_in, a->T_out);
    if (a->layout != 0 && a->layout != 1) return fail(-7, "vae_conv: layout must be 0 (channels-last) or 1 (planes)");
    if (a->layout == 0 && (a->ldy < a->Cout || (a->Cout % 32 == 0 && a->ldy % 8 != 0))) return fail(-8, "vae_conv: bad ldy=%lld", (long long)a->ldy);
    if (a->residual != nullptr && a->ld_res < a->Cout) return fail(-9, "vae_conv: bad ld_res");
    if ((reinterpret_cast<uintptr_t>(a->x) | reinterpret_cast<uintptr_t>(a->w)) & 15)
        return fail(-10, "vae_conv: x and w must be 16-byte aligned");
    if (a->layout == 0 && a->Cout % 32 == 0 && (reinterpret_cast<uintptr_t>(a->y) & 15))
        return fail(-10, "vae_conv: y must be 16-byte aligned");
    if (a->residual != nullptr && a->Cout % 32 == 0 && ((reinterpret_cast<uintptr_t>(a->residual) & 15) || a->ld_res % 8 != 0))
        return fail(-10, "vae_conv: residual must be 16-byte aligned with ld_res %% 8 == 0");
    if (a->stats != nullptr) {
        const int G = a->stat_groups;
        if (a->layout != 0 || G <= 0 || G > CV_MAX_GROUPS || a->Cout % 32 != 0 || a->Cout % G != 0)
            return fail(-11, "vae_conv: epilogue statistics need layout 0, Cout %% 32 == 0 and 1 <= stat_groups <= %d dividing Cout", CV_MAX_GROUPS);
        const int cg = a->Cout / G;
        if (!(cg == 4 || cg == 8 || cg == 16 || cg % 32 == 0))
            return fail(-11, "vae_conv: epilogue statistics need Cout / stat_groups in {4, 8, 16} or a multiple of 32 (got %d)", cg);
        if (reinterpret_cast<uintptr_t>(a->stats) & 7) return fail(-11, "vae_conv: stats must be 8-byte aligned");
    }
    const int s = a->stride_hw;
    CUtensorMap tx, tw;
    {
        // tap-reuse kernel: stride 1, kw = 3, 128-channel weight tiles, at least two waves of pair tiles
        const int64_t t3 = int64_t(a->T_out) * ((a->H_out + C3_TH - 1) / C3_TH) * ((a->W_out + C3_TW - 1) / C3_TW);
        // ... and a width its 128-pixel tiles cover without much overhang: a 180-wide layer (the 120 x 180 level, every tile of
        // the tiled coder at 1/4 resolution) computes 256 columns there but 192 with the 8 x 32 tiles of the kernels below,
        // which more than pays for their three-fold activation traffic (the tap-reuse kernel is ~10 % faster at equal work)
        const int64_t area3 = int64_t((a->W_out + C3_TW - 1) / C3_TW * C3_TW) * ((a->H_out + C3_TH - 1) / C3_TH * C3_TH);
        const int64_t area2 = int64_t((a->W_out + CV_TW - 1) / CV_TW * CV_TW) * ((a->H_out + CV_TH - 1) / CV_TH * CV_TH);
        // narrow outputs (the decoder's 128 -> 3 conv_out): one 32-column tile instead of 64 padded columns
        const bool narrow = a->Cout <= 32;
        const int bn3 = narrow ? 32 : 128;
        const int n_tiles3 = narrow ? 1 : a->Cout_pad / 128;
        if (g_conv_impl == 3 && s == 1 && a->kw == 3 && (narrow || a->Cout_pad % 128 == 0) && ((t3 + 1) / 2) * n_tiles3 >= sm_count() &&
            area3 * 100 <= area2 * 112) {
            const uint64_t dims3[4] = {uint64_t(a->Cin), uint64_t(a->W_in), uint64_t(a->H_in), uint64_t(a->T_in)};
            const uint64_t strides3[3] = {uint64_t(a->Cin) * 2, uint64_t(a->W_in) * a->Cin * 2, uint64_t(a->H_in) * a->W_in * a->Cin * 2};
            const uint32_t box3[4] = {64, uint32_t(C3_TW + 2), uint32_t(C3_TH), 1};
            const uint32_t estr3[4] = {1, 1, 1, 1};
            int rc3 = make_tmap_nd(&tx, a->x, 4, dims3, strides3, box3, estr3);
            if (rc3) return rc3;
            const int K3 = a->kt * a->kh * a->kw * a->Cin;
            rc3 = make_tmap_2d(&tw, a->w, uint64_t(K3), uint64_t(a->Cout_pad), uint64_t(K3) * 2, 64, uint32_t(bn3 / 2));
            if (rc3) return rc3;
            ConvParams p{};
            p.T_out = a->T_out; p.H_out = a->H_out; p.W_out = a->W_out; p.Cout = a->Cout;
            p.kt = a->kt; p.kh = a->kh; p.kw = a->kw; p.stride = 1; p.pad_h = a->pad_h0; p.pad_w = a->pad_w0;
            p.c_chunks = a->Cin / 64;
            p.h_tiles = (a->H_out + C3_TH - 1) / C3_TH;
            p.w_tiles = (a->W_out + C3_TW - 1) / C3_TW;
            p.n_tiles = n_tiles3;
            p.bias = reinterpret_cast<const __nv_bfloat16*>(a->bias);
            p.residual = reinterpret_cast<const __nv_bfloat16*>(a->residual);
            p.ld_res = a->ld_res;
            p.y = reinterpret_cast<__nv_bfloat16*>(a->y);
            p.ldy = a->ldy;
            p.plane_stride = a->plane_stride;
            p.layout = a->layout;
            p.stats = a->stats;
            p.stat_groups = a->stat_groups;
            return narrow ? launch_conv3<32>(tx, tw, p, static_cast<cudaStream_t>(stream))
                          : launch_conv3<128>(tx, tw, p, static_cast<cudaStream_t>(stream));
        }
    }
    const uint64_t dims[4] = {uint64_t(a->Cin), uint64_t(a->W_in), uint64_t(a->H_in), uint64_t(a->T_in)};
    const uint64_t strides[3] = {uint64_t(a->Cin) * 2, uint64_t(a->W_in) * a->Cin * 2, uint64_t(a->H_in) * a->W_in * a->Cin * 2};
    const uint32_t box[4] = {64, uint32_t(CV_TW * s), uint32_t(CV_TH * s), 1};
    const uint32_t estr[4] = {1, uint32_t(s), uint32_t(s), 1};
    int rc = make_tmap_nd(&tx, a->x, 4, dims, strides, box, estr);
    if (rc) return rc;
    const int K = a->kt * a->kh * a->kw * a->Cin;
    // CTA pairs halve the number of work units: keep single-CTA tiles for small layers (tiled coding, low-resolution blocks)
    // where there would be fewer than two waves of pair tiles
    const int64_t px_tiles = int64_t(a->T_out) * ((a->H_out + CV_TH - 1) / CV_TH) * ((a->W_out + CV_TW - 1) / CV_TW);
    const int pair_bn = a->Cout_pad % 256 == 0 ? 256 : 128;
    // ... unless the pair tiles fit ONE round of the machine where the single-CTA tiles need two or more: a pair CTA computes
    // twice the output columns from the same activation box (60 x 90 x 512 layers: 96 CTAs x 1 round instead of 192 tiles on 148)
    const int64_t pair_tiles = ((px_tiles + 1) / 2) * (a->Cout_pad / pair_bn);
    const int64_t rounds1 = (px_tiles * (a->Cout_pad / 128) + sm_count() - 1) / sm_count();
    const int64_t rounds2 = (2 * pair_tiles + sm_count() - 1) / sm_count();
    const bool pair = g_conv_impl >= 2 && a->Cout_pad % 128 == 0 && (pair_tiles >= sm_count() || rounds2 * 4 < rounds1 * 3);
    const int bn = pair ? (a->Cout_pad % 256 == 0 ? 256 : 128) : ((a->Cout_pad % 128 == 0) ? 128 : 64);
    rc = make_tmap_2d(&tw, a->w, uint64_t(K), uint64_t(a->Cout_pad), uint64_t(K) * 2, 64, uint32_t(pair ? bn / 2 : bn));
    if (rc) return rc;
    ConvParams p{};
    p.T_out = a->T_out; p.H_out = a->H_out; p.W_out = a->W_out; p.Cout = a->Cout;
    p.kt = a->kt; p.kh = a->kh; p.kw = a->kw; p.stride = s; p.pad_h = a->pad_h0; p.pad_w = a->pad_w0;
    p.c_chunks = a->Cin / 64;
    p.h_tiles = (a->H_out + CV_TH - 1) / CV_TH;
    p.w_tiles = (a->W_out + CV_TW - 1) / CV_TW;
    p.n_tiles = a->Cout_pad / bn;
    p.bias = reinterpret_cast<const __nv_bfloat16*>(a->bias);
    p.residual = reinterpret_cast<const __nv_bfloat16*>(a->residual);
    p.ld_res = a->ld_res;
    p.y = reinterpret_cast<__nv_bfloat16*>(a->y);
    p.ldy = a->ldy;
    p.plane_stride = a->plane_stride;
    p.layout = a->layout;
    p.stats = a->stats;
    p.stat_groups = a->stat_groups;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (pair) return bn == 256 ? launch_conv2<256>(tx, tw, p, st) : launch_conv2<128>(tx, tw, p, st);
    return bn == 128 ? launch_conv<128>(tx, tw, p, st) : launch_conv<64>(tx, tw, p, st);
}

#ifdef TG_DEVELOPER
extern "C" int tg_set_conv_impl(int impl) {  // developer hook (1 = single-CTA tiles, 2 = CTA pairs); not in the public header
    if (impl < 1 || impl > 3) return fail(-1, "conv impl must be 1, 2 or 3");
    g_conv_impl = impl;
    return 0;
}
#endif
