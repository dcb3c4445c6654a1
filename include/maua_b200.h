/*
 * maua_b200.h — C ABI of libmaua_b200.so: the sm_100a synthesis hot path of the audio-reactive StyleGAN2
 * pipeline (drop-in for the native side of JCBrouwer/maua-stylegan2's `op/` extensions and the
 * `models/stylegan2.py:Generator` forward).
 *
 * Conventions (every entry point):
 *   - plain pointers and sizes only; every data pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - `stream` is a `cudaStream_t` passed as `void*` (NULL = legacy default stream); all work is asynchronous
 *     on that stream, nothing synchronises, nothing allocates device memory (the caller owns every buffer;
 *     reference ownership rule: inputs borrowed, outputs caller-allocated — SURVEY.md §8(b) "Op ABI");
 *   - return value 0 = success, negative = error (MAUA_E_*); never throws across the ABI;
 *     `maua_last_error()` returns a thread-local human-readable message for the last failure.
 *   - all tensors are dense fp32 unless stated; NCHW unless the name says nhwc.
 *
 * Reference interfaces replaced (file:line in /root/reference):
 *   maua_upfirdn2d_f32        <- torch::Tensor upfirdn2d(input,kernel,up_x,up_y,down_x,down_y,pad_x0,pad_x1,pad_y0,pad_y1)
 *                                op/upfirdn2d.cpp:12-23, kernel op/upfirdn2d_kernel.cu:49-207
 *   maua_fused_bias_act_f32   <- torch::Tensor fused_bias_act(input,bias,refer,act,grad,alpha,scale)
 *                                op/fused_bias_act.cpp:11-21, kernel op/fused_bias_act_kernel.cu:18-49
 *   maua_linear_f32           <- EqualLinear.forward                      models/stylegan2.py:140-146
 *   maua_style_prologue_f32   <- truncation lerp + modulation EqualLinear + demod coefficients
 *                                models/stylegan2.py:541-543, :220-225
 *   maua_modconv_simt_f32     <- ModulatedConv2d.forward (fp32 SIMT, exact-order fallback) models/stylegan2.py:217-254
 *   maua_modconv_tc           <- ModulatedConv2d.forward (tcgen05 tensor-core path)       models/stylegan2.py:217-254
 *   maua_blur_act_nhwc        <- Blur.forward + NoiseInjection + FusedLeakyReLU           models/stylegan2.py:89-92,262-266
 *   maua_noise_bias_act_f32   <- NoiseInjection.forward + FusedLeakyReLU.forward models/stylegan2.py:262-266, op/fused_act.py:82-97
 *   maua_torgb_f32            <- ToRGB.forward (1x1 modconv + bias + Upsample(skip))      models/stylegan2.py:356-365
 *   maua_synth_forward        <- Generator.forward (synthesis network, one call per batch)  models/stylegan2.py:537-576
 *   maua_rgb_to_u8_nhwc       <- render.split_batches clamp/scale/permute/astype(uint8)   render.py:40-43
 *   maua_fit_frames_u8        <- 2048-wide frames cropped + PIL-resized to 1920x1080             render.py:98-105
 *   maua_bend_warp_f32        <- Translate / Zoom / Rotate network bends                  audioreactive/bend.py:51-102
 *   maua_perlin_noise         <- perlin_noise                                             audioreactive/latent.py:188-246
 */
#ifndef MAUA_B200_H
#define MAUA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAUA_OK 0
#define MAUA_E_ARG (-1)      /* invalid argument / unsupported shape */
#define MAUA_E_CUDA (-2)     /* CUDA runtime / driver error (launch, tensor-map encode, ...) */
#define MAUA_E_UNSUPPORTED (-3)

#define MAUA_ABI_VERSION 2

int maua_abi_version(void);
const char* maua_last_error(void);
/* Number of kernels launched through this library by the calling process (bench.py's gpu_launches). */
long long maua_launch_count(void);
/* Kernel variant and tile configuration chosen by the calling thread's last maua_modconv_tc call, e.g.
 * "v2 up=0 R=4 BN=32 cat=1 groups=1 AS=2 SA=2 SB=9 grid=148" or "v1 up=1 KC=64 BN=128 S=4 grid=96"
 * (diagnostics / tests: which of the two tensor-core kernels a shape was routed to). */
const char* maua_modconv_tc_last_config(void);

/* ------------------------------------------------------------------------------------------------------------
 * Operator ABI (reference op/ extensions)
 * ---------------------------------------------------------------------------------------------------------- */

/* x: [major, in_h, in_w, minor]  ->  y: [major, out_h, out_w, minor],
 * out = (in*up + pad0 + pad1 - k + down) / down  (op/upfirdn2d_kernel.cu:237-240); k: [kh, kw] (un-flipped).
 * Accumulation order per output is y-major / x-minor fp32 FMA, identical to the reference kernels. */
int maua_upfirdn2d_f32(const float* x, float* y, const float* k, int major, int in_h, int in_w, int minor, int kh,
                       int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0,
                       int pad_y1, void* stream);

/* y[i] = act(x[i] + b[(i / step_b) % size_b]) * scale;  act: 1 linear, 3 leaky-relu(alpha);
 * grad: 0 forward, 1 gated by ref, 2 zeros.  b / ref may be NULL (size_b = 0). */
int maua_fused_bias_act_f32(const float* x, const float* b, const float* ref, float* y, long long n, int step_b,
                            int size_b, int act, int grad, float alpha, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Style / mapping
 * ---------------------------------------------------------------------------------------------------------- */

/* y[b,n] = act( sum_k x[b,k] * w[n,k] * w_scale + bias[n] * bias_scale );  act 0: none, 1: lrelu(0.2)*sqrt(2),
 * act 2: as 1 but bias[0] for every n — what the reference CUDA op computes for the 3-D [1,1,N] tensors of
 * `Generator(map_latents=True)` (models/stylegan2.py:506-509, op/fused_bias_act_kernel.cu:29,67-69).
 * pixel_norm 1 first normalises each row of x: x * rsqrt(mean(x^2) + 1e-8) (models/stylegan2.py:15-20);
 * pixel_norm 2 normalises each ELEMENT by itself (the same module over the singleton axis of [1,1,N]). */
int maua_linear_f32(const float* x, const float* w, const float* bias, float* y, int batch, int in_dim, int out_dim,
                    float w_scale, float bias_scale, int act, int pixel_norm, void* stream);

/* One modulated layer of the style prologue.  All pointers are device pointers. */
typedef struct MauaStyleJob {
  const float* mod_w; /* [cin, style_dim]  conv.modulation.weight */
  const float* mod_b; /* [cin]             conv.modulation.bias   */
  const float* wsq;   /* [cout, cin] = w_scale^2 * sum_k W^2, or NULL when demodulate=False */
  float* s_out;       /* [batch, cin]  */
  float* d_out;       /* [batch, cout] (ignored when wsq == NULL) */
  /* Optional range normalisation for consumers that store x * s in fp16 (the "f16" activation format): when not NULL
   * (and wsq != NULL), s_norm_out[b,:] = s[b,:] * 2^-e_b and d_out[b,:] is multiplied by 2^e_b, with e_b the smallest
   * integer such that max_ci |s[b,ci]| < 2^e_b.  Powers of two: conv(x * s) * d is unchanged bit for bit, but
   * |x * s_norm| <= |x|, so trained checkpoints with large styles cannot push the fp16 operand past 65504. */
  float* s_norm_out;  /* [batch, cin] or NULL */
  int32_t cin;
  int32_t cout;
  int32_t latent_index; /* which W+ row feeds this layer */
  int32_t reserved;
} MauaStyleJob;

/* For every job j and sample b:
 *   w      = mean[:] + psi[b] * (latent[b, jobs[j].latent_index, :] - mean[:])      (psi NULL -> psi_scalar)
 *   s[b,:] = w @ (mod_w / sqrt(style_dim))^T + mod_b
 *   d[b,:] = rsqrt( sum_ci s[b,ci]^2 * wsq[:,ci] + 1e-8 )          (* 2^e_b when s_norm_out is given, see above)
 * `jobs` is a DEVICE array of n_jobs MauaStyleJob.  latent: [batch, n_latent, style_dim]; mean: [style_dim] or NULL
 * (NULL = no truncation).  latent_trunc_out (may be NULL): [batch, n_latent, style_dim] truncated latents. */
int maua_style_prologue_f32(const MauaStyleJob* jobs, int n_jobs, const float* latent, const float* mean,
                            const float* psi, float psi_scalar, float* latent_trunc_out, int batch, int n_latent,
                            int style_dim, void* stream);

/* wsq[co,ci] = w_scale^2 * sum_{ky,kx} w[co,ci,ky,kx]^2   (w: [cout,cin,k,k]) */
int maua_weight_sq_f32(const float* w, float* wsq, int cout, int cin, int ksize, float w_scale, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * fp32 SIMT path (exact-order fallback; also the generic path for arbitrary shapes)
 * ---------------------------------------------------------------------------------------------------------- */

/* y[b,co] = d[b,co] * sum_{ci,ky,kx} w_scale*w[co,ci,ky,kx] * s[b,ci] * x[b,ci, ...]
 *   up == 0: same-resolution, zero padding ksize/2                      -> y [B,Cout,H,W]
 *   up == 1: stride-2 transposed conv, no padding (ksize must be 3)      -> y [B,Cout,2H+1,2W+1]
 * s may be NULL (== 1), d may be NULL (== 1). */
int maua_modconv_simt_f32(const float* x, const float* w, const float* s, const float* d, float* y, int batch,
                          int cin, int cout, int h, int w_, int ksize, int up, float w_scale, void* stream);

/* y = lrelu( (x + noise_weight[0] * noise[b*noise_bstride + pix]) + bias[c], slope ) * scale
 * noise may be NULL; noise_bstride is 0 for a broadcast [1,1,H,W] buffer, H*W for per-sample noise. */
int maua_noise_bias_act_f32(const float* x, const float* noise, const float* noise_weight, const float* bias,
                            float* y, int batch, int ch, int h, int w, long long noise_bstride, float slope,
                            float scale, void* stream);

/* rgb[b,r] = sum_c (w_scale * wrgb[r,c] * s[b,c]) * x[b,c] + bias[r] + upfirdn2d(skip, k4, up=2, pad=(2,1))[b,r]
 * skip: [B,3,H/2,W/2] or NULL;  k4: [4,4] FIR (upsample.kernel buffer). */
int maua_torgb_f32(const float* x, const float* wrgb, const float* s, const float* bias, const float* skip,
                   const float* k4, float* y, int batch, int cin, int h, int w, float w_scale, void* stream);

/* wr[b,k,c] = w_scale * wrgb[k,c] * s[b,c]   (per-sample ToRGB rows for the fused conv epilogue) */
int maua_rgb_weights_f32(const float* wrgb, const float* s, float* wr, int batch, int cin, float w_scale, void* stream);

/* y[b,k] = (partial[b,k] + bias[k]) + upfirdn2d(skip, k4, up=2, pad=(2,1))[b,k]   (ToRGB tail after a fused epilogue) */
int maua_rgb_finish_f32(const float* partial, const float* bias, const float* skip, const float* k4, float* y,
                        int batch, int h, int w, void* stream);

/* maua_rgb_finish_f32 followed by maua_rgb_to_u8_nhwc in one pass (the last ToRGB of a frame when only bytes are wanted):
 * out[b,y,x,k] = u8( (partial[b,k] + bias[k]) + up2(skip)[b,k] ); w % 4 == 0.  The fp32 image is never materialised. */
int maua_rgb_finish_u8(const float* partial, const float* bias, const float* skip, const float* k4, uint8_t* out, int batch,
                       int h, int w, void* stream);

/* out[b,y,x,c] = (uint8) trunc( (clamp(rgb[b,c,y,x], -1, 1) + 1) * 127.5 ) */
int maua_rgb_to_u8_nhwc(const float* rgb, uint8_t* out, int batch, int h, int w, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Tensor-core path (tcgen05 / TMEM / TMA), NHWC split-bf16 activations
 *
 * Activations between layers are stored channels-last as TWO bf16 planes (hi, lo) with hi + lo ~= fp32 value
 * (16 mantissa bits), already multiplied by the consuming layer's style s[b,ci]; weights are packed once as
 * [tap][cout][cin] bf16 hi/lo with w_scale folded in.  The conv evaluates hi*hi + hi*lo + lo*hi on the tensor
 * cores with fp32 accumulation in TMEM (n_products = 3, |err| ~ 2^-16 rel) or hi*hi only (n_products = 1).
 * ---------------------------------------------------------------------------------------------------------- */

/* w [cout,cin,k,k] fp32 -> w_hi / w_lo [k*k][cout][cin] bf16 (value = w * w_scale). */
int maua_pack_weight_bf16x2(const float* w, void* w_hi, void* w_lo, int cout, int cin, int ksize, float w_scale,
                            void* stream);

/* x [B,C,H,W] fp32 (x_bstride = 0 broadcasts one sample, e.g. the constant input) times s[b,c]
 * -> x_hi / x_lo [B,H,W,C] bf16. s may be NULL. */
int maua_modulate_split_nhwc(const float* x, long long x_bstride, const float* s, void* x_hi, void* x_lo, int batch,
                             int ch, int h, int w, void* stream);

/* The "f16" activation format (n_products == 2 of maua_modconv_tc): ONE fp16 plane per activation (11-bit mantissa,
 * saturating conversion) and weights as an fp16 (hi, lo) pair (22 bits: exact for practical purposes), so a modulated
 * conv costs one tensor-core pass over the activations instead of three.  Measured network-level error of rounding the
 * activations of the four >= 512^2 layers of the 1024^2 generator to fp16: 4.8e-4 of the tensor max (tools/exp_precision.py)
 * — inside the 1e-3 parity bar, 10x closer to fp32 than the reference's own default GPU path (TF32 cuDNN convs, 5e-3). */
int maua_pack_weight_f16x2(const float* w, void* w_hi, void* w_lo, int cout, int cin, int ksize, float w_scale,
                           void* stream);
/* x [B,C,H,W] fp32 * s[b,c] -> x_f16 [B,H,W,C] fp16 (one plane). s may be NULL. */
int maua_modulate_f16_nhwc(const float* x, long long x_bstride, const float* s, void* x_f16, int batch, int ch, int h,
                           int w, void* stream);

typedef struct MauaConvEpilogue {
  const float* d;            /* [B,Cout] demod or NULL                                                    */
  const float* noise;        /* [B or 1, H_out, W_out] or NULL                                            */
  const float* noise_weight; /* device scalar                                                             */
  const float* bias;         /* [Cout] or NULL                                                            */
  const float* s_next;       /* [B,Cout] style of the consuming layer (applied before the split) or NULL  */
  void* out_hi;              /* [B,H_out,W_out,Cout] bf16 or NULL                                         */
  void* out_lo;              /*  "                                                                         */
  float* out_f32_nchw;       /* [B,Cout,H_out,W_out] post-activation fp32 or NULL                         */
  float* out_raw_nhwc;       /* up==1 only: [B,2H+1,2W+1,Cout] fp32 = d * convT(x) (pre-blur)             */
  long long noise_bstride;   /* 0 or H_out*W_out                                                          */
  float slope;               /* leaky-relu slope (0.2)                                                    */
  float act_scale;           /* sqrt(2)                                                                   */
  int32_t activate;          /* 0: linear (no noise/bias/act), 1: noise+bias+lrelu                        */
  int32_t out_fmt;           /* format of out_hi/out_lo: 0 = bf16 (hi, lo) pair, 1 = ONE fp16 plane in out_hi      */
  /* Optional fused ToRGB partial sums (same-resolution layers whose whole Cout fits one N tile, Cout <= 128):
   * rgb_out[b,k,y,x] = sum_c rgb_w[b,k,c] * act[b,c,y,x]  (no bias / skip: see maua_rgb_finish_f32).            */
  const float* rgb_w;        /* [B,3,Cout] = w_scale * Wrgb[k,c] * s_rgb[b,c] (maua_rgb_weights_f32) or NULL     */
  float* rgb_out;            /* [B,3,H,W] fp32                                                                    */
  /* Optional scratch for the deterministic split-K of the tiny (<= 32^2) layers: ZERO-initialised once by the caller,
   * restored to that state by every launch, private to one stream.  NULL (or too small) disables split-K.          */
  void* workspace;
  long long workspace_bytes;
} MauaConvEpilogue;

/* 3x3 modulated conv on the tensor cores.  x_hi/x_lo [B,H,W,Cin] bf16 (pre-scaled by s), w_hi/w_lo [9][Cout][Cin].
 *   n_products: 3 = split bf16 (hi*hi + hi*lo + lo*hi), 1 = bf16 hi*hi only (fast, ~1e-2),
 *               2 = fp16: x_hi is ONE fp16 plane (x_lo ignored), w_hi/w_lo the fp16 pair (halo kernel only: H,W >= 64x32).
 *   up == 0: same resolution, zero pad 1, fused epilogue (demod, noise, bias, lrelu, next-style, split)
 *   up == 1: stride-2 transposed conv evaluated as 4 sub-pixel phases; writes ep->out_raw_nhwc (times ep->d if given).
 * Requirements: Cin % 32 == 0, Cout % 16 == 0, Cout >= 16.  `ep` is a HOST pointer (copied at launch). */
int maua_modconv_tc(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo,
                    const MauaConvEpilogue* ep_host, int batch, int cin, int cout, int h, int w, int up,
                    int n_products, void* stream);

/* u [B,Hu,Wu,C] fp32 (Hu = 2H+1) -> 4x4 FIR k4 with pad (1,1) -> [B,Hu-1,Wu-1,C], then * ep->d[b,c] when ep->d is not
 * NULL (demodulation commutes with the per-channel FIR: pass the transposed conv a NULL d and give it here, so that the
 * conv epilogue streams its raw phases from TMEM to HBM), then the activation epilogue of `ep` (noise, bias, lrelu,
 * s_next, split / fp16 plane / fp32 NCHW). */
int maua_blur_act_nhwc(const float* u, const float* k4, const MauaConvEpilogue* ep_host, int batch, int ch, int hu,
                       int wu, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Whole-forward handle API: ONE call per batch = models/stylegan2.py:Generator.forward(styles, noise=..., truncation=...,
 * input_is_latent=True, randomize_noise=False) on the tensor-core path (models/stylegan2.py:537-576), for host languages
 * that bind the library directly (SURVEY.md §8(b) "Op ABI", last cell).  The handle is a host object; it owns no device
 * memory: packed weights live in the caller's `plan` buffer, styles / activations in the caller's `workspace`.
 *
 *   maua_synth_create(desc, &h)                       describe the network (parameter pointers are borrowed, like a module)
 *   maua_synth_prepare(h, plan, plan_bytes, stream)   Wsq + packed tensor-core weights; again after the weights change
 *   maua_synth_bind(h, workspace, bytes, batch, st)   lay out the workspace for one batch size (synchronises; not capturable)
 *   maua_synth_forward(h, ...)                        ~50 kernel launches on `stream`, no host synchronisation, capturable
 *   maua_synth_destroy(h)
 * Not covered (use the per-operator entry points, as maua_stylegan2_b200/synthesis.py does): network bends between
 * layers, returning the activation maps, LatentInput (--noconst), layers outside the tensor-core shape set.
 * ---------------------------------------------------------------------------------------------------------- */

typedef struct MauaSynthLayer {      /* one StyledConv (+ the ToRGB that follows it, if any); all DEVICE pointers */
  const float* conv_weight;          /* conv.weight             [1,Cout,Cin,3,3]                                  */
  const float* mod_weight;           /* conv.modulation.weight  [Cin,style_dim]                                   */
  const float* mod_bias;             /* conv.modulation.bias    [Cin]                                             */
  const float* noise_weight;         /* noise.weight            [1]                                               */
  const float* act_bias;             /* activate.bias           [Cout]                                            */
  const float* noise_buffer;         /* noises.noise_i [1,1,H_out,W_out]: used when no per-frame noise is passed  */
  const float* blur_kernel;          /* conv.blur.kernel [4,4] (up layers), else NULL                             */
  const float* rgb_weight;           /* to_rgb.conv.weight [1,3,Cout,1,1] or NULL when no ToRGB follows           */
  const float* rgb_mod_weight;       /* to_rgb.conv.modulation.weight [Cout,style_dim]                            */
  const float* rgb_mod_bias;         /* to_rgb.conv.modulation.bias   [Cout]                                      */
  const float* rgb_bias;             /* to_rgb.bias [1,3,1,1]                                                     */
  const float* rgb_up_kernel;        /* to_rgb.upsample.kernel [4,4] (NULL for the first ToRGB: no skip)          */
  int32_t cin, cout, up;             /* up: 1 = stride-2 transposed conv + blur (models/stylegan2.py:229-238)     */
  int32_t latent_index;              /* W+ row of this conv                                */
  int32_t rgb_latent_index;          /* W+ row of the ToRGB                                                       */
  int32_t reserved;
} MauaSynthLayer;

typedef struct MauaSynthDesc {
  const MauaSynthLayer* layers;      /* HOST array, copied by maua_synth_create                                   */
  const float* const_input;          /* input.input [1,C,in_h,in_w] (ConstantInput)                               */
  int32_t n_layers, in_h, in_w;      /* 2*log2(size)-3 layers; 4 x 4                                              */
  int32_t style_dim, n_latent;       /* 512; 2*log2(size)-2                                                       */
  int32_t precision;                 /* 0 = bf16x3, 1 = mixed (fp16 activations from f16_min_res on), 2 = bf16    */
  int32_t f16_min_res;               /* 512                                                                       */
  int32_t min_rgb_size;              /* Generator(min_rgb_size=4)                                                 */
} MauaSynthDesc;

typedef struct MauaSynth MauaSynth;

int maua_synth_create(const MauaSynthDesc* desc_host, MauaSynth** out);
void maua_synth_destroy(MauaSynth* h);
size_t maua_synth_plan_bytes(const MauaSynth* h);
int maua_synth_prepare(MauaSynth* h, void* plan, size_t plan_bytes, void* stream);
size_t maua_synth_workspace_bytes(const MauaSynth* h, int batch);
int maua_synth_bind(MauaSynth* h, void* workspace, size_t workspace_bytes, int batch, void* stream);
/* latent [batch, latent_rows >= n_latent, style_dim]; noise: HOST array of n_layers device pointers (entry or array NULL
 * -> the layer's noise_buffer), noise_bstride: HOST array (0 = one map for the batch, H*W = per-sample maps);
 * mean_latent [style_dim] or NULL (no truncation), psi [batch] or NULL (-> psi_scalar).
 * out_rgb [batch,3,H,W] fp32 and/or out_u8 [batch,H,W,3] (render.py:40-43 conversion); at least one. */
int maua_synth_forward(MauaSynth* h, const float* latent, int latent_rows, const float* const* noise_host,
                       const long long* noise_bstride_host, const float* mean_latent, const float* psi, float psi_scalar,
                       int batch, float* out_rgb, uint8_t* out_u8, void* stream);
/* [batch, n_latent, style_dim] truncated latents of the last forward (inside the bound workspace) */
const float* maua_synth_truncated_latents(const MauaSynth* h);

/* ------------------------------------------------------------------------------------------------------------
 * Audio feature chain (cuFFT-fronted; replaces the librosa/scipy CPU path of audioreactive/signal.py:31-156 and the
 * latent glue of audioreactive/latent.py:15-26 + examples/default.py:12-25).  Spectrograms: [n_frames][n_bins]
 * interleaved (re, im) fp32, n_bins = n_fft/2 + 1.  cuFFT plans (and their internal work areas) are cached inside the
 * library per (size, batch, device); everything else is caller-allocated.
 * ---------------------------------------------------------------------------------------------------------- */

/* librosa.stft(y, n_fft, hop, window=hann, center=True, pad_mode=reflect); frames_ws: n_frames*n_fft floats */
int maua_audio_stft_f32(const float* y, long long n, float* spec, float* frames_ws, int n_fft, int hop, int n_frames,
                        void* stream);
/* librosa.istft(spec, hop, window=hann, center=True, length=n); `spec` may be clobbered */
int maua_audio_istft_f32(float* spec, float* y, long long n, float* frames_ws, int n_fft, int hop, int n_frames,
                         void* stream);
/* librosa.decompose.hpss soft mask (median 31x1 / 1x31, `power`, `margin`) applied to spec; which: 0 harmonic, 1 percussive.
 * mag_ws: n_frames*n_bins floats */
int maua_audio_hpss_f32(const float* spec, float* spec_out, float* mag_ws, int n_frames, int n_bins, float margin,
                        float power, int which, void* stream);
/* out[t][m] = sum_f |spec[t][f]|^2 * fb[m][f]  (mel or chroma filterbank, fb dense [n_filters][n_bins]) */
int maua_audio_filterbank_f32(const float* spec, const float* fb, float* out, int n_frames, int n_bins, int n_filters,
                              void* stream);
/* librosa.onset.onset_strength tail: mel (power, overwritten with dB) -> power_to_db(amin, top_db) -> lag-1 positive
 * difference -> mean over bands -> left-pad `pad` frames.  scalar_ws: 1 float */
int maua_audio_onset_env_f32(float* mel, float* env, float* scalar_ws, int n_frames, int n_mels, int pad, float amin,
                             float top_db, void* stream);
/* librosa.feature.rms(S=|spec|) */
/* madmom-flavoured onsets (reference default `type="mm"`, signal.py:52-67).
 * stft_mm: FramedSignal(frame_size n_fft, hop, centred, zero padded) * np.hanning, circular shift n_fft/2, R2C ->
 *          spec [n_frames][n_fft/2+1] (madmom drops the Nyquist bin: pass n_bins = n_fft/2 below).
 * onsets_mm: fb [n_bands][n_bins] (LogarithmicFilterbank, transposed), band_lo/hi [n_bands] = bin range of each band
 *          widened by one neighbour (ComplexFlux mask); onset[t] = spectral_diff + spectral_flux + superflux +
 *          complex_flux + modified_kullback_leibler.  filt_ws: n_frames*n_bands floats, lgd_ws: n_frames*n_bins. */
int maua_audio_stft_mm_f32(const float* y, long long n, float* spec, float* frames_ws, int n_fft, int hop, int n_frames,
                           void* stream);
int maua_audio_onsets_mm_f32(const float* spec, const float* fb, const int* band_lo, const int* band_hi, float* onset,
                             float* filt_ws, float* lgd_ws, int n_frames, int spec_stride, int n_bins, int n_bands,
                             int diff_frames, void* stream);
int maua_audio_rms_f32(const float* spec, float* rms, int n_frames, int n_bins, int n_fft, void* stream);
/* CENS post-processing of a chromagram [n_frames][n_chroma]: L1, quantise, hann(win_len) smoothing, L2. ws: same size */
int maua_audio_cens_f32(const float* raw, float* cens, float* ws, int n_frames, int n_chroma, int win_len, void* stream);
/* np.minimum(x, nn_filter(x, aggregate=median, metric=cosine)) with k nearest frames; scratch: scratch_rows*n_frames floats */
int maua_audio_nn_filter_f32(const float* x, float* out, float* scratch, int scratch_rows, int n_frames, int n_chroma,
                             int k, void* stream);
/* scipy.signal.resample along axis 0 of x[n_in][channels] (double precision inside).
 * ws: (n_in + n_out)*C doubles + (n_in/2 + n_out/2 + 2)*C double2 */
int maua_resample_f32(const float* x, float* y, double* ws, int n_in, int n_out, int channels, void* stream);
/* x = clip(x, min(ref), max(ref))   (signal.py:68,94) */
int maua_clip_to_range_f32(const float* ref, int n_ref, float* x, int n, void* stream);
/* audioreactive.gaussian_filter (signal.py:319-368) along axis 0 of x[n_frames][inner], circular;
 * causal_mode 0: symmetric, 1: future taps * causal, 2: future taps zeroed (non-float `causal`, e.g. 0) */
int maua_gaussian_filter_f32(const float* x, float* y, int n_frames, long long inner, float sigma, float smf,
                             int causal_mode, float causal, void* stream);
/* audioreactive.percentile_clip (signal.py:273-292) followed by ** power */
int maua_percentile_clip_f32(const float* x, float* y, int n, float percentile, float power, void* stream);
/* scipy.signal.sosfilt (double-precision state), sos: [n_sections][6] doubles on the device */
int maua_sosfilt_f32(const float* x, float* y, long long n, const double* sos, int n_sections, void* stream);
/* latent.py:15-26: out[t][e] = sum_n chroma[t][n] * selection[n][e] */
int maua_chroma_weight_latents_f32(const float* chroma, const float* selection, float* out, int n_frames, int n_select,
                                   long long latent_elems, void* stream);
/* examples/default.py:20-21: x[t][e] = envelope[t]*target[e] + (1 - envelope[t])*x[t][e] */
int maua_envelope_blend_f32(float* x, const float* envelope, const float* target, int n_frames, long long inner,
                            void* stream);

/* render.py:98-105 on the device: crop [crop_y0:+crop_h, crop_x0:+crop_w] of uint8 NHWC frames [B,in_h,in_w,3] and
 * resample to [B,out_h,out_w,3] exactly like PIL.Image.resize(..., BILINEAR) (Pillow ImagingResample: horizontal then
 * vertical pass, 22-bit fixed-point coefficients, 8-bit rounding after each pass).  bounds_*: [out][2] (first source
 * index inside the crop, tap count), coef_*: [out][ksize] int32 — DEVICE arrays built by the host from Pillow's
 * precompute_coeffs rule (maua_stylegan2_b200/render.py:pillow_bilinear_coeffs). */
int maua_fit_frames_u8(const uint8_t* in, uint8_t* out, int batch, int in_h, int in_w, int crop_y0, int crop_x0,
                       int crop_h, int crop_w, int out_h, int out_w, const int* bounds_x, const int* coef_x, int ksize_x,
                       const int* bounds_y, const int* coef_y, int ksize_y, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Network bending + looping noise
 * ---------------------------------------------------------------------------------------------------------- */

/* audioreactive/bend.py:51-102 (Translate / Zoom / Rotate) as one kernel:
 *   P   = pad_stages(x) (+ noise)            pad_*_host: n_pad_* (left, right) pairs applied in order, HOST ints;
 *                                            pad_mode 0 = nn.ReflectionPad2d, 1 = nn.ReplicationPad2d
 *   y[b,c,oy,ox] = bilinear_zeros(P[b,c], M_b^-1 (ox + crop_x0, oy + crop_y0))
 * x: [B,C,H,W]; y: [B,C,out_h,out_w]; noise (may be NULL): [noise_b,noise_c,Hp,Wp] with noise_b in {1,B},
 * noise_c in {1,C} (AddNoise, bend.py:28-40); minv: [B,6] row-major 2x3 INVERSE affine (padded-frame pixels,
 * x first): sx = m0*X + m1*Y + m2, sy = m3*X + m4*Y + m5  (kornia warp_affine, bilinear, zeros, align_corners). */
int maua_bend_warp_f32(const float* x, float* y, const float* noise, const float* minv, int batch, int ch, int h, int w,
                       const int* pad_x_host, int n_pad_x, const int* pad_y_host, int n_pad_y, int pad_mode,
                       int noise_b, int noise_c, int out_h, int out_w, int crop_y0, int crop_x0, void* stream);
/* audioreactive/latent.py:188-246 perlin_noise: gradients [r0+1,r1+1,r2+1,3] fp64 (tileable wrap already applied),
 * out [s0,s1,s2] fp64 (out_f64 != 0) or fp32; s_i % r_i == 0; values in [-1, 1] after the reference's `*2 - 1`. */
int maua_perlin_noise(const double* gradients, void* out, int s0, int s1, int s2, int r0, int r1, int r2, int out_f64,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MAUA_B200_H */
