/* tokensgen_b200 — C ABI of the B200-native TokensGen FIFO-denoising hot path.
 *
 * The reference (Vicky0522/TokensGen) is 100 % Python: this path has no existing FFI.  Each entry point
 * below replaces one group of PyTorch library calls on the reference's hot path; the comment above it
 * cites the reference lines it replaces.  The reference-side binding is a ctypes stub (see INTEGRATION.md
 * and tokensgen_b200/_ext.py).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless its name ends in `_host`.
 *   - the caller owns every buffer; the library allocates nothing and keeps no state between calls except a mutex-guarded
 *     cache of CUtensorMap descriptors keyed by (pointer, shape).
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no internal synchronisation,
 *     so every call is CUDA-graph capturable.
 *   - return 0 = ok; negative = argument error (nothing was launched); positive = cudaError_t / CUresult.
 *     tg_last_error() returns a thread-local message for the last non-zero return.
 *   - activations / weights are bf16 (uint16_t storage) row-major; accumulation is fp32.
 *   - "residual stream" X is [B, rows_per_batch, d] with rows ordered [text | video | vip] per batch —
 *     the order CogVideoXPatchEmbed.forward emits (reference longvgen/models/embeddings.py:541-544).
 */
#ifndef TOKENSGEN_B200_H
#define TOKENSGEN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint16_t tg_bf16; /* raw bfloat16 bits */

#define TG_VERSION 1
#define TG_HEAD_DIM 64 /* CogVideoX attention_head_dim; the kernels are specialised for it */

int tg_version(void);
const char* tg_last_error(void);
/* There are no process-wide settings: the shipped library has one fixed kernel configuration.  (A developer build,
 * `python -m tokensgen_b200.build --dev` -> libtokensgen_b200_dev.so, additionally exports tg_set_tuning / tg_set_gemm_impl /
 * tg_set_conv_impl / tg_debug_attn_trace for the A/B tools under tools/; they are not part of this ABI.)
 *
 * Workspaces: NO entry point needs hidden scratch memory.  Every intermediate an operation needs is an explicit argument
 * (`scratch` of tg_time_embedding, the `sums` / `stats` arrays of the GroupNorm statistics, the output tensors), so there is
 * no tg_*_workspace_bytes query: sizes follow from the documented shapes. */

/* Row layout of the residual stream, shared by the fused epilogues:
 * row r of a batch is text if r < n_text, video if r < n_text + n_video (frame = (r - n_text) / hw), else vip. */
typedef struct {
    int rows_per_batch; /* n_text + n_video + n_vip */
    int n_text;
    int n_video;
    int n_vip;
    int hw;     /* video tokens per latent frame (30*45 = 1350) */
    int frames; /* latent frames per window (13) */
    /* Sequence-parallel shard (Ulysses single-clip path): the buffers of a row-local call hold only rows
     * [row0, row0 + rows_local) of every batch, i.e. they are [B, rows_local, *] and local row i is row row0 + i of its
     * batch.  rows_local == 0 (and row0 == 0): unsharded, all rows_per_batch rows. */
    int row0;
    int rows_local;
} tg_rowmap;

#define TG_MAX_PEERS 8 /* GPUs of one NVSwitch domain */

/* One modulation vector source: row (b*frames + f) of a [B*frames, ld] bf16 table, f = 0 unless the row is a
 * video row.  Replaces the `repeat "b f c -> b (f hw) c"` broadcasts of
 * longvgen/models/normalization.py:448-459 (video rows: per-frame; text rows: frame 0) and :482-487 (vip rows). */
typedef struct {
    const tg_bf16* text;  /* may be NULL if n_text == 0 */
    const tg_bf16* video;
    const tg_bf16* vip;   /* may be NULL if n_vip == 0 */
    int64_t ld_text, ld_video, ld_vip; /* row strides in elements */
} tg_modvec;

/* ---------------------------------------------------------------------------------------------------
 * K10  timestep embedding.  Replaces time_proj + time_embedding:
 *   longvgen/models/cogvideox_transformer_3d.py:669-680, longvgen/models/embeddings.py:28-79,953-965.
 * timesteps: fp32 [R].  out_emb / out_silu: bf16 [R, time_dim] (emb and SiLU(emb), the latter being the
 * input of every AdaLN linear, normalization.py:448).  sincos_dim = 3072, flip_sin_to_cos = 1, freq_shift = 0. */
int tg_time_embedding(const float* timesteps, int R, int sincos_dim, int time_dim, int flip_sin_to_cos,
                      float freq_shift, const tg_bf16* w1, const tg_bf16* b1, const tg_bf16* w2,
                      const tg_bf16* b2, tg_bf16* out_emb, tg_bf16* out_silu, tg_bf16* scratch /* [R, time_dim] */,
                      void* stream);

/* ---------------------------------------------------------------------------------------------------
 * K2  LayerNorm(affine, eps) then x*(1+scale)+shift with per-row modulation vectors.
 * Replaces CogVideoXLayerNormZero / CogVideoXVIPLayerNormZero / (norm_final + AdaLayerNorm) elementwise parts:
 *   longvgen/models/normalization.py:457-459, :487, :70-92; cogvideox_transformer_3d.py:736-747.
 * x, out: [B*rows_per_batch, d] (may alias).  ln_w/ln_b: bf16 [d] per segment (text+video share `ln_*`,
 * vip rows use `vip_ln_*`).  If ln2_w != NULL a second affine LayerNorm (eps2) is applied to video rows before the
 * modulation (norm_final followed by norm_out.norm).  Rows of a segment whose shift/scale pointer is NULL are skipped
 * (left untouched in `out`). */
int tg_ln_modulate(const tg_bf16* x, tg_bf16* out, int B, int d, const tg_rowmap* map, const tg_bf16* ln_w,
                   const tg_bf16* ln_b, const tg_bf16* vip_ln_w, const tg_bf16* vip_ln_b, float eps,
                   const tg_bf16* ln2_w, const tg_bf16* ln2_b, float eps2, const tg_modvec* shift,
                   const tg_modvec* scale, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * GEMM family: out = epilogue(A[M,K] @ W[N,K]^T + bias[N]) on tcgen05 tensor cores (TMA-fed, TMEM accumulators).
 * A: bf16 row-major, leading dim lda; W: bf16 [N,K] row-major (the nn.Linear weight layout, untransposed).
 * K % 64 == 0, N % 64 == 0, all base pointers 16-byte aligned, lda % 8 == 0. */
enum { TG_ACT_NONE = 0, TG_ACT_GELU_TANH = 1, TG_ACT_SILU = 2 };

/* K1/K8a/K9/K11: plain Linear (+activation).  Replaces nn.Linear calls at normalization.py:448,482,
 * embeddings.py:516,521,526 (patchify conv as GEMM over tg_patchify rows), diffusers FeedForward net[0]
 * (GELU-tanh; cogvideox_transformer_3d.py:316,322), proj_out (:748). */
int tg_gemm_bias_act(const tg_bf16* A, int64_t lda, const tg_bf16* W, const tg_bf16* bias, tg_bf16* out,
                     int64_t ldo, int M, int N, int K, int act, void* stream);

/* K7/K8b: X[m,:] += gate(m) * (A @ W^T + bias) in place.  Replaces to_out[0] + gated residual
 * (attention_processor.py:2143-2148 + cogvideox_transformer_3d.py:290-293) and FeedForward net[2] + gated
 * residual (:318-324).  M = B*rows_per_batch. */
int tg_gemm_gate_residual(const tg_bf16* A, int64_t lda, const tg_bf16* W, const tg_bf16* bias, tg_bf16* X,
                          int64_t ldx, int B, int N, int K, const tg_rowmap* map, const tg_modvec* gate,
                          void* stream);

/* K3: fused Q/K/V projections with per-head LayerNorm(64) and 3D-RoPE in the epilogue, written head-major.
 * Replaces attention_processor.py:2009-2056 (and :1915-1937 for the plain processor).
 * W: [nproj*H*64, K] = rows of the projections concatenated in `proj` order.  For projection p and head h the
 * epilogue computes y = A@W_p,h^T + bias; if ln_w: y = LN_64(y)*ln_w+ln_b (eps); then RoPE on interleaved pairs
 * (x0,x1)->(x0*c0 - x1*s0, x1*c1 + x0*s1) with cos/sin rows taken from cos_video[r-n_text] for video rows and
 * cos_vip[r-n_text-n_video] for vip rows (NULL table = no RoPE for that segment; text rows never rotate);
 * result stored at out[((b*H + h)*out_rows + r)*64 ..] iff r < out_rows. */
typedef struct {
    tg_bf16* out;    /* [B, H, out_rows, 64] */
    int out_rows;    /* rows of each batch kept for this projection (prefix of the batch's rows) */
    const tg_bf16* ln_w; /* [64] or NULL */
    const tg_bf16* ln_b;
    const float* cos_video; /* [n_video, 64] fp32 or NULL */
    const float* sin_video;
    const float* cos_vip;   /* [n_vip, 64] fp32 or NULL */
    const float* sin_vip;
} tg_qkv_proj;

int tg_qkv_rope_gemm(const tg_bf16* A, int64_t lda, const tg_bf16* W, const tg_bf16* bias, int B, int H, int K,
                     const tg_rowmap* map, const tg_qkv_proj* proj, int nproj, float ln_eps, void* stream);

/* K3 fused with the first Ulysses all-to-all (sequence-sharded rows -> head-sharded Q/K/V): the same GEMM + epilogue
 * over THIS rank's rows (map->row0 / rows_local), but head h of projection p is stored into rank (h / (H/world))'s
 * buffer peer[p][h / (H/world)], laid out [B, H/world, out_rows, 64], at local head h % (H/world) and batch row r —
 * plain 16-byte stores through NVLink peer mappings issued by the epilogue warps while the tensor cores work on the next
 * tile (the reference has no counterpart: its base stage runs on one GPU, pipeline_cogvideox_mp_fifo.py:1186-1305).
 * peer[p][own rank] is the local buffer.  proj[p].out is ignored.  The caller orders this kernel before the consumers on
 * the other ranks (a cross-rank stream barrier). */
typedef struct {
    int world; /* ranks of the sequence-parallel group; H % world == 0, world <= TG_MAX_PEERS */
    tg_bf16* peer[6][TG_MAX_PEERS];
} tg_qkv_scatter;
int tg_qkv_rope_gemm_sp(const tg_bf16* A, int64_t lda, const tg_bf16* W, const tg_bf16* bias, int B, int H, int K,
                        const tg_rowmap* map, const tg_qkv_proj* proj, int nproj, float ln_eps,
                        const tg_qkv_scatter* scatter, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * K4/K5/K6: non-causal softmax(Q K^T * scale) V, head_dim 64, tcgen05 flash attention.
 * Replaces the three F.scaled_dot_product_attention calls at attention_processor.py:2066-2069, 2117-2125
 * (and :1939 for the plain processor).
 * q: [B,H,*,64] with q_rows queries starting at row q_row0 of a tensor with q_rows_alloc rows per head; k, v likewise.
 * out: [B, out_rows_alloc, H*64] token-major; query i is written to row out_row0 + i.
 * accumulate != 0: out = out + out_scale * attn (the `hidden + scale * text_video_hidden` of :2134), else out = attn. */
int tg_attn_fwd(const tg_bf16* q, int64_t q_rows_alloc, int64_t q_row0, int q_rows, const tg_bf16* k,
                const tg_bf16* v, int64_t kv_rows_alloc, int64_t kv_row0, int kv_rows, tg_bf16* out,
                int64_t out_rows_alloc, int64_t out_row0, int B, int H, float softmax_scale, int accumulate,
                float out_scale, void* stream);

/* K4 + K5 in ONE launch: out[q] = softmax(Q K^T * scale) V  +  out_scale2 * softmax(Q2 K2'^T * scale) V2' for the same
 * q_rows query rows, where K2'/V2' are rows [kv_row0_2, kv_row0_2 + kv_rows2) of k2/v2.  This is the self-attention over
 * [text; video] plus `hidden + scale * cross-attention to the vip tokens` of VideoIPAdapterCogVideoXAttnProcessor2_0
 * (attention_processor.py:2066-2069 followed by :2117-2119,2126-2134): the second, short problem (480 keys) reuses the
 * CTA's TMEM / barrier / K-V ring set-up and its result is added before the output row leaves the SM's cache.
 * q,k,v: [B,H,rows_alloc,64] (rows 0..q_rows / 0..kv_rows used); q2,k2,v2: [B,H,rows_alloc2,64]; out: [B,out_rows_alloc,H*64]. */
int tg_attn_fwd_pair(const tg_bf16* q, const tg_bf16* k, const tg_bf16* v, int64_t rows_alloc, int q_rows, int kv_rows,
                     const tg_bf16* q2, const tg_bf16* k2, const tg_bf16* v2, int64_t rows_alloc2, int64_t kv_row0_2,
                     int kv_rows2, tg_bf16* out, int64_t out_rows_alloc, int B, int H, float softmax_scale, float out_scale2,
                     void* stream);

/* K4/K5/K6 fused with the second Ulysses all-to-all (head-sharded attention output -> sequence-sharded rows): this rank
 * attends over its H local heads (global heads head0 .. head0+H-1) and ALL rows; output row g (= out_row0 + query index) of
 * a batch goes to its owner rank o = min(g / chunk, world-1), whose buffer peer[o] is [B, rows_local(o), H_total*64]
 * (rows_local(o) = chunk, or rows_per_batch - (world-1)*chunk for the last rank), at local row g - o*chunk and columns
 * [(head0 + h)*64, +64) — peer stores (and, with accumulate, peer loads) from the attention epilogue. */
typedef struct {
    int world;
    int chunk;          /* rows of each batch owned by every rank but the last */
    int rows_per_batch; /* rows of the full residual stream per batch */
    int H_total;        /* heads of the model (row stride of the output = H_total*64) */
    int head0;          /* first global head of this call's q/k/v */
    tg_bf16* peer[TG_MAX_PEERS];
} tg_attn_scatter;
int tg_attn_fwd_sp(const tg_bf16* q, int64_t q_rows_alloc, int64_t q_row0, int q_rows, const tg_bf16* k,
                   const tg_bf16* v, int64_t kv_rows_alloc, int64_t kv_row0, int kv_rows, const tg_attn_scatter* out,
                   int64_t out_row0, int B, int H, float softmax_scale, int accumulate, float out_scale, void* stream);
int tg_attn_fwd_pair_sp(const tg_bf16* q, const tg_bf16* k, const tg_bf16* v, int64_t rows_alloc, int q_rows, int kv_rows,
                        const tg_bf16* q2, const tg_bf16* k2, const tg_bf16* v2, int64_t rows_alloc2, int64_t kv_row0_2,
                        int kv_rows2, const tg_attn_scatter* out, int B, int H, float softmax_scale, float out_scale2,
                        void* stream);

/* ---------------------------------------------------------------------------------------------------
 * K9/K11 index maps (bit-exact class).
 * tg_patchify: latents [B,F,C,H,W] -> rows [B*F*(H/p)*(W/p), C*p*p] in Conv2d weight order (c, pi, pj);
 *   replaces the im2col implied by embeddings.py:524-529.
 * tg_unpatchify: rows [B*F*(H/p)*(W/p), C*p*p] -> [B,F,C,H,W]; replaces cogvideox_transformer_3d.py:754-759. */
int tg_patchify(const tg_bf16* latents, tg_bf16* rows, int B, int F, int C, int H, int W, int p, void* stream);
int tg_unpatchify(const tg_bf16* rows, tg_bf16* latents, int B, int F, int C, int H, int W, int p, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * K12: classifier-free guidance + per-frame DPM-Solver++(2M) SDE step for one window, one launch.
 * Replaces cogvideo_sampling_mp_fifo.py:527-550 + scheduling_dpm_cogvideox.py:424-463 (13 Python iterations of
 * ~15 elementwise launches each), and pipeline_cogvideox_mp_fifo.py:1247-1290 for the base stage.
 *   noise_pred : bf16 [n_branches, F, chw]; n_branches = 2 -> (uncond, cond) combined as u + g*(c-u); 1 -> used as is;
 *                3 (BF16_CHAIN only) -> `use_separate_guidance` (:528-530): (uncond_txt, uncond_img, txt_img) combined as
 *                a + guidance_scale * (a - ut) + guidance_scale2 * (a - ui) with a = txt_img, where guidance_scale /
 *                guidance_scale2 are the reference's (guidance_scale - 1) / (guidance_scale_img - 1).
 *   coef       : fp32 [F, 8] per frame {sqrt_alpha_t, sqrt_beta_t, mult0, mult1, mult2, mult3, mult_noise, flags};
 *                flags (as float) 1.0 = second-order frame (old x0 valid and prev_timestep >= 0), else 0.0.
 *                The host computes them in fp64 exactly as CogVideoXDPMScheduler.get_variables/get_mult and casts.
 *   noise1     : the draw a first-order frame uses; noise2: the draw a second-order frame uses (the reference draws
 *                twice on that path and uses the second, scheduling_dpm_cogvideox.py:450,461).  Both bf16 [F, chw].
 *   mode TG_DPM_BF16_CHAIN : every tensor op rounds to bf16 like the reference FIFO worker (all-bf16 tensors);
 *        TG_DPM_BASE_CHAIN : the base pipeline's mixed chain (noise_pred.float(), fp32 x0 history, bf16 latents);
 *   old_x0 / x0_out are bf16 in BF16_CHAIN mode, fp32 (`*_f32`) in BASE_CHAIN mode.
 * With identical noise the result is bit-identical to the reference chain. */
enum { TG_DPM_BF16_CHAIN = 1, TG_DPM_BASE_CHAIN = 0 };
typedef struct {
    const tg_bf16* noise_pred;
    const float* noise_pred_f32; /* BASE_CHAIN only: an already guided fp32 model output [F, chw] (then noise_pred = NULL) */
    int n_branches;
    float guidance_scale;
    const tg_bf16* sample;     /* [F, chw] */
    const tg_bf16* old_x0;     /* [F, chw] or NULL */
    const float* old_x0_f32;   /* [F, chw] or NULL */
    const tg_bf16* noise1;
    const tg_bf16* noise2;
    const float* coef;
    tg_bf16* prev_sample;      /* [F, chw] */
    tg_bf16* x0_out;           /* [F, chw] or NULL */
    float* x0_out_f32;         /* [F, chw] or NULL */
    int F;
    int64_t chw;
    int mode;
    float guidance_scale2;     /* n_branches = 3 only */
} tg_dpm_step_args;
int tg_cfg_dpm_step(const tg_dpm_step_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------------
 * K13: advance the FIFO queue by one latent frame and re-noise the new tail slot, in place.
 * Replaces shift_latents() of cogvideo_sampling_mp_fifo.py:117-131 (clone + slice copy) and
 * CogVideoXDPMScheduler.add_noise_to_xt (scheduling_dpm_cogvideox.py:497-518).
 *   queue : bf16 [n_slots, chw]; slot i <- slot i+1 for i < n_slots-1;
 *           tail <- bf16( sqrt_one_minus_beta * old_tail + sqrt_beta * noise ) evaluated in fp64 like the reference
 *           (its [1]-shaped fp64 coefficients promote the expression to fp64 before the in-place assignment rounds it).
 *   x0_queue : optional bf16 [n_slots, chw] history shifted the same way (tail left untouched: the host marks it empty).
 *   noise : bf16 [chw]. */
int tg_queue_shift_renoise(tg_bf16* queue, tg_bf16* x0_queue, int n_slots, int64_t chw, const tg_bf16* noise,
                           double sqrt_one_minus_beta, double sqrt_beta, void* stream);

/* ===================================================================================================
 * 3D causal VAE (CogVideoX AutoencoderKL) — activations are CHANNELS-LAST bf16 [T, H, W, C] inside the coder (batch 1;
 * the pipeline enables slicing, longvgen infer_cogvideo_mp_fifo.py:181-182).  The host mirror converts at the boundary.
 * ---------------------------------------------------------------------------------------------------
 * K15/K17: implicit-GEMM convolution on tcgen05.  Replaces CogVideoXCausalConv3d.forward -> CogVideoXSafeConv3d
 * (longvgen/models/autoencoder_kl_cogvideox.py:38-64,133-145: torch.cat of the conv cache + F.pad + cuDNN conv3d) and the
 * per-frame Conv2d of diffusers CogVideoXUpsample3D / CogVideoXDownsample3D (called at :413,:606).
 *   x : [T_in, H_in, W_in, Cin], Cin % 64 == 0 (zero-pad narrow inputs).  The first kt-1 frames of x ARE the causal
 *       context (previous call's last frames = the reference's conv_cache, or copies of the first frame, :120-127);
 *       output frame t reads input frames t .. t+kt-1.  H/W zero padding is implicit: pad_h0/pad_w0 rows/cols before
 *       the first pixel, whatever the output extent needs after the last.
 *   w : [Cout_pad, kt*kh*kw*Cin] bf16, tap-major (kt, kh, kw, c) — the nn.Conv3d weight [Cout, Cin, kt, kh, kw] permuted
 *       once by the host; rows >= Cout are zero.  bias [Cout] or NULL.
 *   y : layout 0 = channels-last [T_out, H_out, W_out, ldy]; layout 1 = channel planes, element (n, t, h, w) at
 *       y[n*plane_stride + (t*H_out + h)*W_out + w] (lets conv_out write straight into a [C, T, H, W] video / moments
 *       tensor).  residual (channels-last, row stride ld_res) is added when not NULL (ResnetBlock3D skip, :308). */
typedef struct {
    const tg_bf16* x;
    int T_in, H_in, W_in, Cin;
    const tg_bf16* w;
    const tg_bf16* bias;
    int Cout, Cout_pad;
    int kt, kh, kw;
    int stride_hw;
    int pad_h0, pad_w0;
    int T_out, H_out, W_out;
    const tg_bf16* residual;
    int64_t ld_res;
    tg_bf16* y;
    int64_t ldy;
    int64_t plane_stride;
    int layout;
    /* GroupNorm statistics of the OUTPUT accumulated in the epilogue (layout 0 only): stats[g] += sum, stats[stat_groups + g] +=
     * sum of squares of the stored (bf16-rounded) outputs of group g — the `sums` tg_vae_norm_act reads, so the statistics pass
     * over the tensor (tg_vae_group_stats) is not needed for tensors a convolution produced.  The caller zeroes `stats`
     * (2 * stat_groups doubles); NULL = off.  Needs Cout % 32 == 0, stat_groups <= 64 and Cout / stat_groups in {4, 8, 16} or a
     * multiple of 32. */
    double* stats;
    int stat_groups;
} tg_conv_args;
int tg_vae_conv(const tg_conv_args* args, void* stream);

/* K16: GroupNorm statistics and apply.  Replaces nn.GroupNorm (:241-242,700), CogVideoXSpatialNorm3D.forward (:175-188:
 * F.interpolate of zq + two 1x1x1 convolutions + 3 elementwise ops) and the SiLU that always follows (:290,301,741,879).
 *   tg_vae_group_stats: sums[0..G) += sum x, sums[G..2G) += sum x^2 per group over [pixels, C] (caller zeroes `sums`).
 *   tg_vae_norm_act   : y = act( ((x - mean_g) * rstd_g * gamma + beta) [* zy[z(p)] + zb[z(p)]] ); zy / zb are
 *       conv_y(zq) / conv_b(zq) evaluated at LATENT resolution [Tz, Hz, Wz, C] (a 1x1 conv commutes with nearest
 *       up-sampling) and z(p) is F.interpolate's nearest source pixel, the first frame handled apart when T is odd (:176-184). */
int tg_vae_group_stats(const tg_bf16* x, int64_t pixels, int C, int64_t ldx, int groups, double* sums, void* stream);
typedef struct {
    const tg_bf16* x;
    int64_t ldx;
    int T, H, W, C, groups;
    float eps;
    const double* sums;
    const tg_bf16* gamma;
    const tg_bf16* beta;
    const tg_bf16* zy;
    const tg_bf16* zb;
    int Tz, Hz, Wz;
    int silu;
    tg_bf16* y;
    int64_t ldy;
    int64_t ldz; /* elements between latent pixels of zy / zb (0 = C): lets every SpatialNorm of the decoder read its columns of ONE
                    [latent pixels, sum of 2C] table produced by a single GEMM over all conv_y / conv_b weights */
} tg_norm_args;
int tg_vae_norm_act(const tg_norm_args* args, void* stream);

/* K17: nearest up-sampling of diffusers CogVideoXUpsample3D (mode 0: x2 in H,W per frame; mode 1 = compress_time: x2 in T too,
 * the first frame only in H,W when T is odd and > 1; T == 1 stays 1) and the temporal average pooling of
 * CogVideoXDownsample3D (frame pairs; first frame kept when T is odd).  Channels-last in and out. */
int tg_vae_upsample(const tg_bf16* x, tg_bf16* y, int T, int H, int W, int C, int mode, void* stream);
int tg_vae_avgpool_time(const tg_bf16* x, tg_bf16* y, int T, int64_t frame_elems, void* stream);

/* [C, T, H, W] planes (plane stride in elements) -> channels-last [T*H*W, Cpad], zero pad channels. */
int tg_vae_to_channels_last(const tg_bf16* x, tg_bf16* y, int C, int Cpad, int64_t pixels, int64_t plane_stride, void* stream);

/* K19: z = (mean + exp(0.5 * clamp(logvar, -30, 20)) * eps) * scale; moments = [mean | logvar], n elements each.
 * Replaces DiagonalGaussianDistribution.sample() * scaling_factor (longvgen/pipeline/pipeline_cogvideox_mp_fifo.py:585). */
int tg_vae_posterior_sample(const tg_bf16* moments, const tg_bf16* eps, tg_bf16* z, int64_t n, float scale, void* stream);

/* K18: blend_v (axis 0) / blend_h (axis 1) of tiled_encode / tiled_decode (:1190-1204), in place on b, bf16 op-by-op
 * rounding like the reference.  a: [planes, Ha, Wa], b: [planes, Hb, Wb]; extent is clamped to the tile sizes. */
int tg_vae_blend(const tg_bf16* a, tg_bf16* b, int64_t planes, int Ha, int Wa, int Hb, int Wb, int extent, int axis, void* stream);

/* K20: decoded video planes [3, F, H, W] bf16 in [-1, 1] -> packed frames [F, H, W, 3] uint8,
 * u = rint(clamp(x/2 + 0.5, 0, 1) * 255) in fp32 (round-half-even).  Replaces VideoProcessor.postprocess_video
 * (longvgen/pipeline/pipeline_cogvideox_mp_fifo.py:363: denormalise, permute, float32 numpy frames) plus the exporter's
 * float -> uint8 conversion on the host: 4x fewer bytes cross PCIe and no host float math.
 * pixels = F*H*W; plane_stride = elements between channel planes. */
int tg_vae_frames_to_rgb8(const tg_bf16* x, uint8_t* y, int64_t pixels, int64_t plane_stride, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TOKENSGEN_B200_H */
