// Style prologue: truncation lerp + all modulation affines + demodulation coefficients in ONE launch,
// plus the small EqualLinear used by the mapping network.
//
// Reference: models/stylegan2.py:541-543 (truncation), :140-146 / :207 (modulation EqualLinear, lr_mul=1,
// bias_init=1), :220-225 (modulate / demodulate).  The reference materialises the [B,Cout,Cin,k,k] weight
// tensor per batch; here  d[b,co] = rsqrt( sum_ci s[b,ci]^2 * Wsq[co,ci] + 1e-8 ),  Wsq = c^2 * sum_k W^2
// (SURVEY.md Appendix B.1), so nothing per-sample is ever written to HBM except s and d.
#include "common.cuh"

namespace maua {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Two launches per batch (all 26 modulated layers each):
//   style_s_kernel: grid (n_jobs, ceil(max_cin/64), ceil(B/8)); block = 64 rows of one modulation matrix x 8 samples.
//                   The truncated latents of the 8 samples sit in shared memory, every weight row is read once per
//                   block and reused for the 8 samples (the first version re-read 1 MB of weights per (job, sample)).
//   style_d_kernel: same tiling over Wsq rows for the demodulation vector (needs the complete s of the layer).
constexpr int SB = 8;      // samples per block
constexpr int SROWS = 64;  // weight rows per block (8 per warp)

__global__ void __launch_bounds__(256) style_s_kernel(const MauaStyleJob* __restrict__ jobs,
                                                      const float* __restrict__ latent, const float* __restrict__ mean,
                                                      const float* __restrict__ psi, float psi_scalar,
                                                      float* __restrict__ latent_trunc_out, int batch, int n_latent,
                                                      int style_dim) {
  extern __shared__ float sw[];  // [SB][style_dim]
  const MauaStyleJob job = jobs[blockIdx.x];
  const int row0 = blockIdx.y * SROWS;
  if (row0 >= job.cin) return;
  const int b0 = blockIdx.z * SB;
  const int nb = min(SB, batch - b0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < nb * style_dim; i += 256) {
    const int bl = i / style_dim, k = i - bl * style_dim;
    const int b = b0 + bl;
    float v = latent[((long long)b * n_latent + job.latent_index) * style_dim + k];
    if (mean) {
      const float m = __ldg(mean + k);
      const float t = psi ? __ldg(psi + b) : psi_scalar;
      v = __fadd_rn(m, __fmul_rn(t, __fsub_rn(v, m)));  // w_bar + psi * (w - w_bar), models/stylegan2.py:541-543
    }
    sw[i] = v;
    if (latent_trunc_out && blockIdx.y == 0)
      latent_trunc_out[((long long)b * n_latent + job.latent_index) * style_dim + k] = v;
  }
  __syncthreads();
  const float lin_scale = rsqrtf((float)style_dim);
  for (int r = row0 + warp; r < min(row0 + SROWS, job.cin); r += 8) {
    const float* wr = job.mod_w + (long long)r * style_dim;
    float acc[SB];
#pragma unroll
    for (int j = 0; j < SB; ++j) acc[j] = 0.f;
    for (int k = lane; k < style_dim; k += 32) {
      const float wv = __ldg(wr + k);
#pragma unroll
      for (int j = 0; j < SB; ++j) acc[j] = fmaf(sw[j * style_dim + k], wv, acc[j]);  // rows j >= nb read stale smem: unused
    }
#pragma unroll
    for (int j = 0; j < SB; ++j) acc[j] = warp_sum(acc[j]);
    if (lane < nb) {
      float a = acc[0];
#pragma unroll
      for (int j = 1; j < SB; ++j) a = (lane == j) ? acc[j] : a;
      job.s_out[(long long)(b0 + lane) * job.cin + r] = fmaf(a, lin_scale, __ldg(job.mod_b + r));
    }
  }
}

__global__ void __launch_bounds__(256) style_d_kernel(const MauaStyleJob* __restrict__ jobs, int batch, int max_cin) {
  extern __shared__ float ss[];  // [SB][cin] squared styles
  const MauaStyleJob job = jobs[blockIdx.x];
  const int row0 = blockIdx.y * SROWS;
  if (job.wsq == nullptr || row0 >= job.cout) return;
  const int b0 = blockIdx.z * SB;
  const int nb = min(SB, batch - b0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < SB * job.cin; i += 256) {
    const int bl = i / job.cin, k = i - bl * job.cin;
    float v = 0.f;
    if (bl < nb) v = job.s_out[(long long)(b0 + bl) * job.cin + k];
    ss[i] = v * v;
  }
  __syncthreads();
  // Optional power-of-two range normalisation (s_norm_out): every block of the job holds the complete s^2 rows, so each
  // derives the per-sample exponent itself (no cross-block reduction); block row 0 writes the normalised styles.
  __shared__ int s_exp[SB];
  if (tid < SB) s_exp[tid] = 0;
  if (job.s_norm_out != nullptr) {   // (uniform over the block)
    __syncthreads();
    if (warp < SB) {
      float m = 0.f;
      for (int k = lane; k < job.cin; k += 32) m = fmaxf(m, ss[warp * job.cin + k]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      int e = 0;
      if (m > 0.f) frexpf(sqrtf(m) * 1.0000002f, &e);   // max|s| = f * 2^e, f in [0.5, 1)  ->  |s| * 2^-e < 1
      if (lane == 0) s_exp[warp] = e;
    }
    __syncthreads();
    if (blockIdx.y == 0) {
      for (int i = tid; i < nb * job.cin; i += 256) {
        const int bl = i / job.cin, k = i - bl * job.cin;
        const long long o = (long long)(b0 + bl) * job.cin + k;
        job.s_norm_out[o] = ldexpf(job.s_out[o], -s_exp[bl]);
      }
    }
  } else {
    __syncthreads();
  }
  for (int r = row0 + warp; r < min(row0 + SROWS, job.cout); r += 8) {
    const float* wr = job.wsq + (long long)r * job.cin;
    float acc[SB];
#pragma unroll
    for (int j = 0; j < SB; ++j) acc[j] = 0.f;
    for (int k = lane; k < job.cin; k += 32) {
      const float wv = __ldg(wr + k);
#pragma unroll
      for (int j = 0; j < SB; ++j) acc[j] = fmaf(ss[j * job.cin + k], wv, acc[j]);
    }
#pragma unroll
    for (int j = 0; j < SB; ++j) acc[j] = warp_sum(acc[j]);
    if (lane < nb) {
      float a = acc[0];
#pragma unroll
      for (int j = 1; j < SB; ++j) a = (lane == j) ? acc[j] : a;
      job.d_out[(long long)(b0 + lane) * job.cout + r] = ldexpf(rsqrtf(a + 1e-8f), s_exp[lane]);
    }
  }
}

// grid (ceil(out_dim/8), batch): one warp per output feature.
__global__ void __launch_bounds__(256) linear_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                     const float* __restrict__ bias, float* __restrict__ y,
                                                     int in_dim, int out_dim, float w_scale, float bias_scale,
                                                     int act, int pixel_norm) {
  extern __shared__ float sx[];
  __shared__ float red[8];
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* xr = x + (long long)b * in_dim;
  float sq = 0.f;
  for (int i = tid; i < in_dim; i += 256) {
    const float v = xr[i];
    sx[i] = v;
    sq = fmaf(v, v, sq);
  }
  float norm = 1.f;
  if (pixel_norm == 2) {
    // PixelNorm over a SINGLETON axis (reference map_latents path feeds [1,1,512]: models/stylegan2.py:20,507):
    // every element is normalised by itself, x * rsqrt(x^2 + 1e-8)
    __syncthreads();
    for (int i = tid; i < in_dim; i += 256) {
      const float v = sx[i];
      sx[i] = v * rsqrtf(v * v + 1e-8f);
    }
    __syncthreads();
  } else if (pixel_norm) {
    sq = warp_sum(sq);
    if (lane == 0) red[warp] = sq;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[i];
    norm = rsqrtf(tot / (float)in_dim + 1e-8f);
  } else {
    __syncthreads();
  }
  const int n = blockIdx.x * 8 + warp;
  if (n >= out_dim) return;
  const float* wr = w + (long long)n * in_dim;
  float acc = 0.f;
  for (int i = lane; i < in_dim; i += 32) acc = fmaf(sx[i] * norm, __ldg(wr + i), acc);
  acc = warp_sum(acc);
  if (lane == 0) {
    float v = acc * w_scale;
    // act 2: the reference CUDA op on a 3-D [1,1,N] input indexes bias[(i / step_b) % size_b] with step_b = N, i.e.
    // bias[0] for every feature (op/fused_bias_act_kernel.cu:29,67-69)
    if (bias) v += __ldg(bias + (act == 2 ? 0 : n)) * bias_scale;
    if (act >= 1) v = (v > 0.f ? v : v * 0.2f) * 1.4142135623730951f;
    y[(long long)b * out_dim + n] = v;
  }
}

__global__ void __launch_bounds__(256) weight_sq_kernel(const float* __restrict__ w, float* __restrict__ wsq,
                                                        long long n, int kk, float s2) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float* p = w + i * kk;
    float acc = 0.f;
    for (int t = 0; t < kk; ++t) acc = fmaf(p[t], p[t], acc);
    wsq[i] = acc * s2;
  }
}

}  // namespace maua

extern "C" int maua_style_prologue_f32(const MauaStyleJob* jobs, int n_jobs, const float* latent, const float* mean,
                                       const float* psi, float psi_scalar, float* latent_trunc_out, int batch,
                                       int n_latent, int style_dim, void* stream) {
  using namespace maua;
  MAUA_CHECK_ARG(jobs && latent && n_jobs >= 0 && batch >= 0, "style_prologue: bad arguments");
  MAUA_CHECK_ARG(style_dim >= 1 && style_dim <= 1024, "style_prologue: style_dim must be in [1,1024]");
  if (n_jobs == 0 || batch == 0) return MAUA_OK;
  const int max_dim = 1024;  // upper bound on cin / cout of a job (StyleGAN2: 512); rows beyond a job's size exit early
  const int zb = ceil_div(batch, SB);
  MAUA_CHECK_ARG(zb <= 65535, "style_prologue: batch too large");
  cudaStream_t st = as_stream(stream);
  const size_t smem_s = (size_t)SB * style_dim * sizeof(float);
  const size_t smem_d = (size_t)SB * max_dim * sizeof(float);
  MAUA_CHECK_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(style_s_kernel), 64 * 1024));
  MAUA_CHECK_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(style_d_kernel), 64 * 1024));
  style_s_kernel<<<dim3(n_jobs, max_dim / SROWS, zb), 256, smem_s, st>>>(jobs, latent, mean, psi, psi_scalar,
                                                                        latent_trunc_out, batch, n_latent, style_dim);
  MAUA_CHECK_LAUNCH("style_prologue(s)");
  style_d_kernel<<<dim3(n_jobs, max_dim / SROWS, zb), 256, smem_d, st>>>(jobs, batch, max_dim);
  MAUA_CHECK_LAUNCH("style_prologue(d)");
  return MAUA_OK;
}

extern "C" int maua_linear_f32(const float* x, const float* w, const float* bias, float* y, int batch, int in_dim,
                               int out_dim, float w_scale, float bias_scale, int act, int pixel_norm, void* stream) {
  using namespace maua;
  MAUA_CHECK_ARG(x && w && y && batch >= 0 && in_dim >= 1 && out_dim >= 1, "linear: bad arguments");
  MAUA_CHECK_ARG(in_dim <= 8192, "linear: in_dim > 8192 unsupported");
  if (batch == 0) return MAUA_OK;
  for (int b0 = 0; b0 < batch; b0 += 65535) {
    const int nb = (batch - b0 < 65535) ? batch - b0 : 65535;
    linear_kernel<<<dim3(ceil_div(out_dim, 8), nb), 256, in_dim * sizeof(float), as_stream(stream)>>>(
        x + (long long)b0 * in_dim, w, bias, y + (long long)b0 * out_dim, in_dim, out_dim, w_scale, bias_scale, act,
        pixel_norm);
    MAUA_CHECK_LAUNCH("linear");
  }
  return MAUA_OK;
}

extern "C" int maua_weight_sq_f32(const float* w, float* wsq, int cout, int cin, int ksize, float w_scale,
                                  void* stream) {
  using namespace maua;
  MAUA_CHECK_ARG(w && wsq && cout >= 1 && cin >= 1 && ksize >= 1, "weight_sq: bad arguments");
  const long long n = (long long)cout * cin;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  weight_sq_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(w, wsq, n, ksize * ksize, w_scale * w_scale);
  MAUA_CHECK_LAUNCH("weight_sq");
  return MAUA_OK;
}
