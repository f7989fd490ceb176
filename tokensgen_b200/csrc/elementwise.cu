// HBM-bound and tiny kernels of the DiT window step: LayerNorm+modulate, timestep
// embedding, patchify / unpatchify index maps, fused CFG + DPM-Solver++ step.  All are coalesced 16-byte
// vector kernels; none uses tensor cores (nothing here is a dense contraction worth them).
#include <cuda_bf16.h>

#include "common.h"
#include "ptx.cuh"

namespace tg {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
    f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 v;
    v.x = pack_bf16x2(f[0], f[1]); v.y = pack_bf16x2(f[2], f[3]);
    v.z = pack_bf16x2(f[4], f[5]); v.w = pack_bf16x2(f[6], f[7]);
    return v;
}
__device__ __forceinline__ void unpack8p(const uint4& v, uint64_t (&p)[4]) {   // eight bf16 -> four packed fp32 pairs
    p[0] = pack_f32x2(bf16_lo(v.x), bf16_hi(v.x));
    p[1] = pack_f32x2(bf16_lo(v.y), bf16_hi(v.y));
    p[2] = pack_f32x2(bf16_lo(v.z), bf16_hi(v.z));
    p[3] = pack_f32x2(bf16_lo(v.w), bf16_hi(v.w));
}
__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// ------------------------------------------------------------------------------------------------ K2
struct LnParams {
    const __nv_bfloat16* x;
    __nv_bfloat16* out;
    int M, d;
    tg_rowmap map;
    const __nv_bfloat16 *ln_w, *ln_b, *vip_ln_w, *vip_ln_b;
    float eps;
    const __nv_bfloat16 *ln2_w, *ln2_b;
    float eps2;
    tg_modvec shift, scale;
};

__device__ __forceinline__ const __nv_bfloat16* mod_row(const tg_modvec& v, const tg_rowmap& m, int b, int seg, int frame) {
    const tg_bf16* p;
    if (seg == 0)
        p = v.text ? v.text + int64_t(b * m.frames) * v.ld_text : nullptr;
    else if (seg == 1)
        p = v.video ? v.video + int64_t(b * m.frames + frame) * v.ld_video : nullptr;
    else
        p = v.vip ? v.vip + int64_t(b * m.frames) * v.ld_vip : nullptr;
    return reinterpret_cast<const __nv_bfloat16*>(p);
}

// Common case (no second LayerNorm): the row stays in registers as PACKED bf16 (NV uint4 per lane) and is unpacked in each
// of the three passes (sum, centred sum of squares, normalise + modulate).  Half the registers of the fp32-resident form
// below -> two to three CTAs per SM instead of one, i.e. enough loads in flight to stream HBM.
template <int NV>
__global__ void __launch_bounds__(256, (NV <= 12) ? 2 : 1) ln_modulate_packed_kernel(const __grid_constant__ LnParams p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + warp;
    if (row >= p.M) return;
    const int b = row / p.map.rows_local;  // host entry normalises rows_local (> 0) / row0: sequence-parallel shard
    const int r = row - b * p.map.rows_local + p.map.row0;
    int seg, frame = 0;
    if (r < p.map.n_text) seg = 0;
    else if (r < p.map.n_text + p.map.n_video) { seg = 1; frame = (r - p.map.n_text) / p.map.hw; }
    else seg = 2;
    const __nv_bfloat16* sh = mod_row(p.shift, p.map, b, seg, frame);
    const __nv_bfloat16* sc = mod_row(p.scale, p.map, b, seg, frame);
    if (sh == nullptr || sc == nullptr) return;
    const uint4* w = reinterpret_cast<const uint4*>((seg == 2) ? p.vip_ln_w : p.ln_w) + lane;
    const uint4* bb = reinterpret_cast<const uint4*>((seg == 2) ? p.vip_ln_b : p.ln_b) + lane;
    const uint4* shv4 = reinterpret_cast<const uint4*>(sh) + lane;
    const uint4* scv4 = reinterpret_cast<const uint4*>(sc) + lane;
    const uint4* xr = reinterpret_cast<const uint4*>(p.x + int64_t(row) * p.d) + lane;
    uint4 raw[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) raw[i] = __ldg(xr + i * 32);   // all loads of the row in flight at once
    // All arithmetic on packed fp32 pairs (FADD2 / FFMA2 / FMUL2: one issue slot per two results) — at HBM speed the scalar
    // form needs ~150 issue slots per 16-byte vector and the kernel was bound by them (0.55 of the copy peak in the step).
    // four independent accumulator pairs: with 16 warps per SM a single dependent FADD2 chain (48 links per pass) is latency-,
    // not issue-bound
    uint64_t acc[4] = {0, 0, 0, 0};                             // (+0, +0) each
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        uint64_t f[4];
        unpack8p(raw[i], f);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = add_f32x2(acc[j], f[j]);
    }
    uint64_t tot = add_f32x2(add_f32x2(acc[0], acc[1]), add_f32x2(acc[2], acc[3]));
    const float inv_d = 1.0f / float(p.d);
    const float mean = warp_sum(f32x2_lo(tot) + f32x2_hi(tot)) * inv_d;
    const uint64_t neg_mean = pack_f32x2(-mean, -mean);
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = 0;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        uint64_t f[4];
        unpack8p(raw[i], f);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint64_t dlt = add_f32x2(f[j], neg_mean);
            acc[j] = fma_f32x2(dlt, dlt, acc[j]);
        }
    }
    tot = add_f32x2(add_f32x2(acc[0], acc[1]), add_f32x2(acc[2], acc[3]));
    const float rstd = rsqrtf(warp_sum(f32x2_lo(tot) + f32x2_hi(tot)) * inv_d + p.eps);
    const uint64_t rstd2 = pack_f32x2(rstd, rstd), one2 = pack_f32x2(1.0f, 1.0f);
    uint4* orow = reinterpret_cast<uint4*>(p.out + int64_t(row) * p.d) + lane;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        uint64_t f[4], wv[4], bv[4], shv[4], scv[4];
        unpack8p(raw[i], f);
        unpack8p(__ldg(w + i * 32), wv);
        unpack8p(__ldg(bb + i * 32), bv);
        unpack8p(__ldg(shv4 + i * 32), shv);
        unpack8p(__ldg(scv4 + i * 32), scv);
        uint4 o;
        uint32_t* ow = &o.x;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // fmaf(fmaf((x - mean) * rstd, w, b), 1 + scale, shift), lane by lane
            const uint64_t n = mul_f32x2(add_f32x2(f[j], neg_mean), rstd2);
            const uint64_t y = fma_f32x2(fma_f32x2(n, wv[j], bv[j]), add_f32x2(scv[j], one2), shv[j]);
            ow[j] = pack_bf16x2(f32x2_lo(y), f32x2_hi(y));
        }
        orow[i * 32] = o;
    }
}

// One warp per row; the row stays in registers between the statistics passes (two-pass mean/variance, fp32).
template <int NV>  // uint4 (8 x bf16) vectors per lane: d = NV * 256
__global__ void __launch_bounds__(256) ln_modulate_kernel(const __grid_constant__ LnParams p) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + warp;
    if (row >= p.M) return;
    const int b = row / p.map.rows_local;  // host entry normalises rows_local (> 0) / row0: sequence-parallel shard
    const int r = row - b * p.map.rows_local + p.map.row0;
    int seg, frame = 0;
    if (r < p.map.n_text) seg = 0;
    else if (r < p.map.n_text + p.map.n_video) { seg = 1; frame = (r - p.map.n_text) / p.map.hw; }
    else seg = 2;
    const __nv_bfloat16* sh = mod_row(p.shift, p.map, b, seg, frame);
    const __nv_bfloat16* sc = mod_row(p.scale, p.map, b, seg, frame);
    if (sh == nullptr || sc == nullptr) return;
    const __nv_bfloat16* w = (seg == 2) ? p.vip_ln_w : p.ln_w;
    const __nv_bfloat16* bb = (seg == 2) ? p.vip_ln_b : p.ln_b;

    const uint4* xr = reinterpret_cast<const uint4*>(p.x + int64_t(row) * p.d);
    float v[NV][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const uint4 raw = __ldg(xr + i * 32 + lane);
        unpack8(raw, v[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += v[i][j];
    }
    const float inv_d = 1.0f / float(p.d);
    float mean = warp_sum(s) * inv_d;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float dlt = v[i][j] - mean;
            ss = fmaf(dlt, dlt, ss);
        }
    float rstd = rsqrtf(warp_sum(ss) * inv_d + p.eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float wv[8], bv[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(w) + i * 32 + lane), wv);
        unpack8(__ldg(reinterpret_cast<const uint4*>(bb) + i * 32 + lane), bv);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = fmaf((v[i][j] - mean) * rstd, wv[j], bv[j]);
    }
    if (p.ln2_w != nullptr) {
        s = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) s += v[i][j];
        mean = warp_sum(s) * inv_d;
        ss = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float dlt = v[i][j] - mean;
                ss = fmaf(dlt, dlt, ss);
            }
        rstd = rsqrtf(warp_sum(ss) * inv_d + p.eps2);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            float wv[8], bv[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(p.ln2_w) + i * 32 + lane), wv);
            unpack8(__ldg(reinterpret_cast<const uint4*>(p.ln2_b) + i * 32 + lane), bv);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[i][j] = fmaf((v[i][j] - mean) * rstd, wv[j], bv[j]);
        }
    }
    uint4* orow = reinterpret_cast<uint4*>(p.out + int64_t(row) * p.d);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float shv[8], scv[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(sh) + i * 32 + lane), shv);
        unpack8(__ldg(reinterpret_cast<const uint4*>(sc) + i * 32 + lane), scv);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = fmaf(v[i][j], 1.0f + scv[j], shv[j]);
        orow[i * 32 + lane] = pack8(v[i]);
    }
}

// ------------------------------------------------------------------------------------------------ K10
// h[r, n] = SiLU(bf16(sum_k temb[r,k] * W1[n,k] + b1[n])), temb = bf16(sinusoid(t_r)) built in shared memory.
__global__ void __launch_bounds__(256)
time_embed_l1_kernel(const float* __restrict__ timesteps, int R, int sincos_dim, int time_dim, int flip, float freq_shift,
                     const __nv_bfloat16* __restrict__ w1, const __nv_bfloat16* __restrict__ b1,
                     __nv_bfloat16* __restrict__ scratch) {
    extern __shared__ __nv_bfloat16 temb[];  // [sincos_dim]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * 8 + warp;
    const int half = sincos_dim / 2;
    const float neg_log = -9.210340371976184f;  // -ln(10000), as torch: fp32(-math.log(10000)) * arange
    for (int r = 0; r < R; ++r) {
        const float t = timesteps[r];
        for (int i = threadIdx.x; i < half; i += blockDim.x) {
            const float e = __fdiv_rn(__fmul_rn(neg_log, float(i)), float(half) - freq_shift);
            const float arg = __fmul_rn(t, expf(e));
            const float sv = sinf(arg), cv = cosf(arg);
            // reference: cat[sin, cos], then flip_sin_to_cos swaps the halves -> [cos, sin]
            temb[i] = __float2bfloat16_rn(flip ? cv : sv);
            temb[half + i] = __float2bfloat16_rn(flip ? sv : cv);
        }
        __syncthreads();
        if (n < time_dim) {
            float acc = 0.f;
            const __nv_bfloat16* wr = w1 + int64_t(n) * sincos_dim;
            for (int k = lane * 8; k < sincos_dim; k += 256) {
                float wv[8], xv[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(wr + k)), wv);
                unpack8(*reinterpret_cast<const uint4*>(temb + k), xv);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc = fmaf(wv[j], xv[j], acc);
            }
            acc = warp_sum(acc);
            if (lane == 0) {
                const float h = round_bf16(acc + __bfloat162float(b1[n]));
                scratch[int64_t(r) * time_dim + n] = __float2bfloat16_rn(h / (1.0f + expf(-h)));
            }
        }
        __syncthreads();
    }
}
// emb[r, n] = sum_k h[r,k] * W2[n,k] + b2[n];  out_silu = SiLU(emb)
__global__ void __launch_bounds__(256)
time_embed_l2_kernel(const __nv_bfloat16* __restrict__ h, int R, int time_dim, const __nv_bfloat16* __restrict__ w2,
                     const __nv_bfloat16* __restrict__ b2, __nv_bfloat16* __restrict__ out_emb,
                     __nv_bfloat16* __restrict__ out_silu) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * 8 + warp;
    if (n >= time_dim) return;
    const __nv_bfloat16* wr = w2 + int64_t(n) * time_dim;
    for (int r = 0; r < R; ++r) {
        float acc = 0.f;
        for (int k = lane * 8; k < time_dim; k += 256) {
            float wv[8], xv[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(wr + k)), wv);
            unpack8(__ldg(reinterpret_cast<const uint4*>(h + int64_t(r) * time_dim + k)), xv);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc = fmaf(wv[j], xv[j], acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            const float e = round_bf16(acc + __bfloat162float(b2[n]));
            out_emb[int64_t(r) * time_dim + n] = __float2bfloat16_rn(e);
            out_silu[int64_t(r) * time_dim + n] = __float2bfloat16_rn(e / (1.0f + expf(-e)));
        }
    }
}

// ------------------------------------------------------------------------------------------------ K9 / K11
// rows[((b*F+f)*Hp + hh)*Wp + ww][(c*p + pi)*p + pj] <-> latents[b][f][c][hh*p+pi][ww*p+pj]
template <bool TO_ROWS>
__global__ void patch_map_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int BF, int C,
                                 int H, int W, int p) {
    const int Hp = H / p, Wp = W / p;
    const int64_t total = int64_t(BF) * C * H * W;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        // i enumerates the latent layout [bf][c][y][x] (coalesced on that side)
        const int x = int(i % W);
        const int y = int((i / W) % H);
        const int c = int((i / (int64_t(W) * H)) % C);
        const int bf = int(i / (int64_t(W) * H * C));
        const int hh = y / p, pi = y - hh * p, ww = x / p, pj = x - ww * p;
        const int64_t ridx = ((int64_t(bf) * Hp + hh) * Wp + ww) * (C * p * p) + (c * p + pi) * p + pj;
        if (TO_ROWS) dst[ridx] = src[i];
        else dst[i] = src[ridx];
    }
}

// ------------------------------------------------------------------------------------------------ K12
struct DpmParams {
    tg_dpm_step_args a;
};

// Rounding helpers: `rb` rounds an fp32 op result to bf16 (what a bf16 tensor op stores), fp32 ops use _rn intrinsics
// so the compiler cannot contract them into FMAs (torch evaluates each op separately).
__device__ __forceinline__ float rb(float x) { return round_bf16(x); }

__global__ void __launch_bounds__(256) cfg_dpm_step_kernel(const __grid_constant__ tg_dpm_step_args a) {
    const int f = blockIdx.y;
    const float* cf = a.coef + f * 8;
    const float sa = cf[0], sb = cf[1], m0 = cf[2], m1 = cf[3], m2 = cf[4], m3 = cf[5], mn = cf[6];
    const bool second = cf[7] != 0.0f;
    const int64_t nvec = a.chw / 8;
    const int64_t frame_off = int64_t(f) * a.chw;
    const int64_t branch_stride = int64_t(a.F) * a.chw;
    for (int64_t vi = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; vi < nvec; vi += int64_t(gridDim.x) * blockDim.x) {
        const int64_t off = frame_off + vi * 8;
        float u[8], c[8], c3[8], smp[8], nz[8], old[8], mo[8], x0[8], ps[8];
        if (a.noise_pred_f32 != nullptr) {
            const float4 o0 = *reinterpret_cast<const float4*>(a.noise_pred_f32 + off);
            const float4 o1 = *reinterpret_cast<const float4*>(a.noise_pred_f32 + off + 4);
            u[0] = o0.x; u[1] = o0.y; u[2] = o0.z; u[3] = o0.w;
            u[4] = o1.x; u[5] = o1.y; u[6] = o1.z; u[7] = o1.w;
        } else {
            unpack8(*reinterpret_cast<const uint4*>(a.noise_pred + off), u);
        }
        unpack8(*reinterpret_cast<const uint4*>(a.sample + off), smp);
        unpack8(*reinterpret_cast<const uint4*>((second ? a.noise2 : a.noise1) + off), nz);
        if (a.n_branches >= 2) unpack8(*reinterpret_cast<const uint4*>(a.noise_pred + branch_stride + off), c);
        if (a.n_branches == 3) unpack8(*reinterpret_cast<const uint4*>(a.noise_pred + 2 * branch_stride + off), c3);
        if (second) {
            if (a.mode == TG_DPM_BF16_CHAIN) {
                unpack8(*reinterpret_cast<const uint4*>(a.old_x0 + off), old);
            } else {
                const float4 o0 = *reinterpret_cast<const float4*>(a.old_x0_f32 + off);
                const float4 o1 = *reinterpret_cast<const float4*>(a.old_x0_f32 + off + 4);
                old[0] = o0.x; old[1] = o0.y; old[2] = o0.z; old[3] = o0.w;
                old[4] = o1.x; old[5] = o1.y; old[6] = o1.z; old[7] = o1.w;
            }
        }
        if (a.mode == TG_DPM_BF16_CHAIN) {
            // every tensor is bf16: each op = fp32 compute on bf16 inputs, rounded to bf16
            // (cogvideo_sampling_mp_fifo.py:531-533; scheduling_dpm_cogvideox.py:439-463)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (a.n_branches == 3) {
                    // use_separate_guidance (cogvideo_sampling_mp_fifo.py:528-530): u = uncond_txt, c = uncond_img, c3 = txt_img;
                    // txt_img + (g - 1) * (txt_img - uncond_txt) + (g_img - 1) * (txt_img - uncond_img), left to right
                    const float t1 = rb(c3[j] + rb(a.guidance_scale * rb(c3[j] - u[j])));
                    mo[j] = rb(t1 + rb(a.guidance_scale2 * rb(c3[j] - c[j])));
                } else {
                    mo[j] = (a.n_branches == 2) ? rb(u[j] + rb(a.guidance_scale * rb(c[j] - u[j]))) : u[j];
                }
                x0[j] = rb(rb(sa * smp[j]) - rb(sb * mo[j]));
                const float mz = rb(mn * nz[j]);
                if (!second) {
                    ps[j] = rb(rb(rb(m0 * smp[j]) - rb(m1 * x0[j])) + mz);
                } else {
                    const float dd = rb(rb(m2 * x0[j]) - rb(m3 * old[j]));
                    ps[j] = rb(rb(rb(m0 * smp[j]) - rb(m1 * dd)) + mz);
                }
            }
        } else {
            // base stage: noise_pred.float() -> fp32 CFG; x0 and its history fp32; latents/noise bf16
            // (pipeline_cogvideox_mp_fifo.py:1247,1262-1263,1280-1290)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                mo[j] = (a.n_branches == 2) ? __fadd_rn(u[j], __fmul_rn(a.guidance_scale, __fsub_rn(c[j], u[j]))) : u[j];
                x0[j] = __fsub_rn(rb(sa * smp[j]), __fmul_rn(sb, mo[j]));
                const float mz = rb(mn * nz[j]);
                if (!second) {
                    ps[j] = __fadd_rn(__fsub_rn(rb(m0 * smp[j]), __fmul_rn(m1, x0[j])), mz);
                } else {
                    const float dd = __fsub_rn(__fmul_rn(m2, x0[j]), __fmul_rn(m3, old[j]));
                    ps[j] = __fadd_rn(__fsub_rn(rb(m0 * smp[j]), __fmul_rn(m1, dd)), mz);
                }
            }
        }
        *reinterpret_cast<uint4*>(a.prev_sample + off) = pack8(ps);
        if (a.x0_out != nullptr) *reinterpret_cast<uint4*>(a.x0_out + off) = pack8(x0);
        if (a.x0_out_f32 != nullptr) {
            *reinterpret_cast<float4*>(a.x0_out_f32 + off) = make_float4(x0[0], x0[1], x0[2], x0[3]);
            *reinterpret_cast<float4*>(a.x0_out_f32 + off + 4) = make_float4(x0[4], x0[5], x0[6], x0[7]);
        }
    }
}

// ------------------------------------------------------------------------------------------------ K13
// One thread owns one 16-byte column of every slot and walks the slots in order, so the in-place shift has no
// cross-thread hazard.
__global__ void __launch_bounds__(256)
queue_shift_renoise_kernel(__nv_bfloat16* __restrict__ queue, __nv_bfloat16* __restrict__ x0q, int n_slots, int64_t chw,
                           const __nv_bfloat16* __restrict__ noise, double s1, double s2) {
    const int64_t nvec = chw / 8;
    for (int64_t vi = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; vi < nvec; vi += int64_t(gridDim.x) * blockDim.x) {
        uint4* q = reinterpret_cast<uint4*>(queue) + vi;
        for (int s = 0; s + 1 < n_slots; ++s) q[int64_t(s) * nvec] = q[int64_t(s + 1) * nvec];
        if (x0q != nullptr) {
            uint4* h = reinterpret_cast<uint4*>(x0q) + vi;
            for (int s = 0; s + 1 < n_slots; ++s) h[int64_t(s) * nvec] = h[int64_t(s + 1) * nvec];
        }
        float x[8], n[8], o[8];
        unpack8(q[int64_t(n_slots - 1) * nvec], x);
        unpack8(reinterpret_cast<const uint4*>(noise)[vi], n);
        // fp64 -> bf16 in one rounding (no float intermediate), like the reference's fp64 expression assigned into bf16
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const double dv = __dadd_rn(__dmul_rn(s1, double(x[j])), __dmul_rn(s2, double(n[j])));
            o[j] = __bfloat162float(__double2bfloat16(dv));
        }
        q[int64_t(n_slots - 1) * nvec] = pack8(o);
    }
}

}  // namespace tg

using namespace tg;

extern "C" int tg_queue_shift_renoise(tg_bf16* queue, tg_bf16* x0_queue, int n_slots, int64_t chw, const tg_bf16* noise,
                                      double sqrt_one_minus_beta, double sqrt_beta, void* stream) {
    if (!queue || !noise) return fail(-1, "queue_shift_renoise: null pointer");
    if (n_slots < 1 || chw <= 0 || chw % 8 != 0) return fail(-2, "queue_shift_renoise: n_slots=%d chw=%lld", n_slots, (long long)chw);
    const int64_t nvec = chw / 8;
    const int grid = int((nvec + 255) / 256);
    queue_shift_renoise_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<__nv_bfloat16*>(queue), reinterpret_cast<__nv_bfloat16*>(x0_queue), n_slots, chw,
        reinterpret_cast<const __nv_bfloat16*>(noise), sqrt_one_minus_beta, sqrt_beta);
    return check_launch("queue_shift_renoise");
}

extern "C" int tg_ln_modulate(const tg_bf16* x, tg_bf16* out, int B, int d, const tg_rowmap* map, const tg_bf16* ln_w,
                              const tg_bf16* ln_b, const tg_bf16* vip_ln_w, const tg_bf16* vip_ln_b, float eps,
                              const tg_bf16* ln2_w, const tg_bf16* ln2_b, float eps2, const tg_modvec* shift,
                              const tg_modvec* scale, void* stream) {
    if (!x || !out || !map || !shift || !scale) return fail(-1, "ln_modulate: null pointer");
    if (B <= 0 || d <= 0 || d % 256 != 0) return fail(-2, "ln_modulate: d=%d must be a positive multiple of 256", d);
    if (map->rows_per_batch != map->n_text + map->n_video + map->n_vip || map->rows_per_batch <= 0 || map->frames <= 0 ||
        (map->n_video > 0 && (map->hw <= 0 || map->n_video != map->hw * map->frames)))
        return fail(-3, "ln_modulate: inconsistent rowmap");
    if (!rowmap_shard_ok(map)) return fail(-3, "ln_modulate: rowmap shard [row0, row0+rows_local) outside the batch");
    if (!ln_w || !ln_b) return fail(-4, "ln_modulate: ln_w/ln_b required");
    if (map->n_vip > 0 && shift->vip && (!vip_ln_w || !vip_ln_b)) return fail(-5, "ln_modulate: vip rows need vip_ln_w/b");
    if ((ln2_w == nullptr) != (ln2_b == nullptr)) return fail(-6, "ln_modulate: ln2_w/ln2_b must come together");
    LnParams p{};
    p.x = reinterpret_cast<const __nv_bfloat16*>(x);
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    p.map = normalised_rowmap(map);
    p.M = B * p.map.rows_local;
    p.d = d;
    p.ln_w = reinterpret_cast<const __nv_bfloat16*>(ln_w);
    p.ln_b = reinterpret_cast<const __nv_bfloat16*>(ln_b);
    p.vip_ln_w = reinterpret_cast<const __nv_bfloat16*>(vip_ln_w);
    p.vip_ln_b = reinterpret_cast<const __nv_bfloat16*>(vip_ln_b);
    p.eps = eps;
    p.ln2_w = reinterpret_cast<const __nv_bfloat16*>(ln2_w);
    p.ln2_b = reinterpret_cast<const __nv_bfloat16*>(ln2_b);
    p.eps2 = eps2;
    p.shift = *shift;
    p.scale = *scale;
    const int grid = (p.M + 7) / 8;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (ln2_w == nullptr) {
        switch (d / 256) {
            case 1: ln_modulate_packed_kernel<1><<<grid, 256, 0, st>>>(p); break;
            case 2: ln_modulate_packed_kernel<2><<<grid, 256, 0, st>>>(p); break;
            case 4: ln_modulate_packed_kernel<4><<<grid, 256, 0, st>>>(p); break;
            case 8: ln_modulate_packed_kernel<8><<<grid, 256, 0, st>>>(p); break;
            case 12: ln_modulate_packed_kernel<12><<<grid, 256, 0, st>>>(p); break;
            case 16: ln_modulate_packed_kernel<16><<<grid, 256, 0, st>>>(p); break;
            default: return fail(-7, "ln_modulate: d=%d not instantiated (256, 512, 1024, 2048, 3072, 4096)", d);
        }
        return check_launch("ln_modulate");
    }
    switch (d / 256) {
        case 1: ln_modulate_kernel<1><<<grid, 256, 0, st>>>(p); break;
        case 2: ln_modulate_kernel<2><<<grid, 256, 0, st>>>(p); break;
        case 4: ln_modulate_kernel<4><<<grid, 256, 0, st>>>(p); break;
        case 8: ln_modulate_kernel<8><<<grid, 256, 0, st>>>(p); break;
        case 12: ln_modulate_kernel<12><<<grid, 256, 0, st>>>(p); break;
        case 16: ln_modulate_kernel<16><<<grid, 256, 0, st>>>(p); break;
        default: return fail(-7, "ln_modulate: d=%d not instantiated (256, 512, 1024, 2048, 3072, 4096)", d);
    }
    return check_launch("ln_modulate");
}

extern "C" int tg_time_embedding(const float* timesteps, int R, int sincos_dim, int time_dim, int flip_sin_to_cos,
                                 float freq_shift, const tg_bf16* w1, const tg_bf16* b1, const tg_bf16* w2,
                                 const tg_bf16* b2, tg_bf16* out_emb, tg_bf16* out_silu, tg_bf16* scratch, void* stream) {
    if (!timesteps || !w1 || !b1 || !w2 || !b2 || !out_emb || !out_silu || !scratch)
        return fail(-1, "time_embedding: null pointer");
    if (R <= 0 || sincos_dim <= 0 || time_dim <= 0 || sincos_dim % 16 != 0 || time_dim % 8 != 0)
        return fail(-2, "time_embedding: bad dims R=%d sincos_dim=%d time_dim=%d", R, sincos_dim, time_dim);
    if (sincos_dim * 2 > 48 * 1024) return fail(-3, "time_embedding: sincos_dim too large for shared memory");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int grid = (time_dim + 7) / 8;
    time_embed_l1_kernel<<<grid, 256, sincos_dim * 2, st>>>(
        timesteps, R, sincos_dim, time_dim, flip_sin_to_cos, freq_shift, reinterpret_cast<const __nv_bfloat16*>(w1),
        reinterpret_cast<const __nv_bfloat16*>(b1), reinterpret_cast<__nv_bfloat16*>(scratch));
    int rc = check_launch("time_embedding/l1");
    if (rc) return rc;
    time_embed_l2_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(scratch), R, time_dim,
                                               reinterpret_cast<const __nv_bfloat16*>(w2),
                                               reinterpret_cast<const __nv_bfloat16*>(b2),
                                               reinterpret_cast<__nv_bfloat16*>(out_emb),
                                               reinterpret_cast<__nv_bfloat16*>(out_silu));
    return check_launch("time_embedding/l2");
}

static int patch_common(const void* a, const void* b, int B, int F, int C, int H, int W, int p) {
    if (!a || !b) return fail(-1, "patchify: null pointer");
    if (B <= 0 || F <= 0 || C <= 0 || H <= 0 || W <= 0 || p <= 0 || H % p != 0 || W % p != 0)
        return fail(-2, "patchify: bad dims B=%d F=%d C=%d H=%d W=%d p=%d", B, F, C, H, W, p);
    return 0;
}
extern "C" int tg_patchify(const tg_bf16* latents, tg_bf16* rows, int B, int F, int C, int H, int W, int p, void* stream) {
    int rc = patch_common(latents, rows, B, F, C, H, W, p);
    if (rc) return rc;
    const int64_t total = int64_t(B) * F * C * H * W;
    const int grid = int((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    patch_map_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(latents), reinterpret_cast<__nv_bfloat16*>(rows), B * F, C, H, W, p);
    return check_launch("patchify");
}
extern "C" int tg_unpatchify(const tg_bf16* rows, tg_bf16* latents, int B, int F, int C, int H, int W, int p, void* stream) {
    int rc = patch_common(rows, latents, B, F, C, H, W, p);
    if (rc) return rc;
    const int64_t total = int64_t(B) * F * C * H * W;
    const int grid = int((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    patch_map_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const __nv_bfloat16*>(rows), reinterpret_cast<__nv_bfloat16*>(latents), B * F, C, H, W, p);
    return check_launch("unpatchify");
}

extern "C" int tg_cfg_dpm_step(const tg_dpm_step_args* a, void* stream) {
    if (!a) return fail(-1, "cfg_dpm_step: null args");
    if ((!a->noise_pred && !a->noise_pred_f32) || !a->sample || !a->noise1 || !a->noise2 || !a->coef || !a->prev_sample)
        return fail(-1, "cfg_dpm_step: null pointer");
    if (a->noise_pred_f32 && (a->noise_pred || a->n_branches != 1 || a->mode != TG_DPM_BASE_CHAIN))
        return fail(-6, "cfg_dpm_step: noise_pred_f32 is for the base chain with n_branches = 1 and no bf16 noise_pred");
    if (a->n_branches < 1 || a->n_branches > 3) return fail(-2, "cfg_dpm_step: n_branches must be 1, 2 or 3");
    if (a->n_branches == 3 && a->mode != TG_DPM_BF16_CHAIN)
        return fail(-2, "cfg_dpm_step: three guidance branches are the FIFO worker's bf16 chain only (the base stage guides in fp32 on the host side)");
    if (a->F <= 0 || a->chw <= 0 || a->chw % 8 != 0) return fail(-3, "cfg_dpm_step: F=%d chw=%lld (chw %% 8 == 0)", a->F, (long long)a->chw);
    if (a->mode == TG_DPM_BF16_CHAIN) {
        if (a->old_x0_f32 || a->x0_out_f32) return fail(-4, "cfg_dpm_step: bf16 chain takes bf16 x0 buffers");
    } else if (a->mode == TG_DPM_BASE_CHAIN) {
        if (a->old_x0 || a->x0_out) return fail(-4, "cfg_dpm_step: base chain takes fp32 x0 buffers");
    } else {
        return fail(-5, "cfg_dpm_step: unknown mode %d", a->mode);
    }
    // A second-order frame needs its history; the flags live on the device, so require the buffer whenever any could.
    const int64_t nvec = a->chw / 8;
    dim3 grid(unsigned((nvec + 255) / 256), unsigned(a->F));
    cfg_dpm_step_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(*a);
    return check_launch("cfg_dpm_step");
}
