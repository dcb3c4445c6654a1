// Elementwise bias / noise / leaky-ReLU kernels (HBM-bound, 8 B per element).
//
//   maua_fused_bias_act_f32  — drop-in for op/fused_bias_act_kernel.cu:18-49 (all act/grad codes).
//   maua_noise_bias_act_f32  — NoiseInjection (models/stylegan2.py:262-266) + FusedLeakyReLU
//                              (op/fused_act.py:82-97) in one pass; rounding mirrors the reference's two
//                              elementwise ops exactly: t = x + (w*noise); t = t + bias; y = lrelu(t)*scale.
#include "common.cuh"

namespace maua {

__device__ __forceinline__ float bias_act_one(float x, float ref, int code, float alpha, float scale) {
  float yv;
  switch (code) {
    case 30: yv = (x > 0.f) ? x : __fmul_rn(x, alpha); break;
    case 31: yv = (ref > 0.f) ? x : __fmul_rn(x, alpha); break;
    case 12:
    case 32: yv = 0.f; break;
    default: yv = x; break;  // 10, 11 and the reference's `default`
  }
  return __fmul_rn(yv, scale);
}

// Vector path: step_b % 4 == 0 (so a float4 never straddles two bias entries) and 16-byte aligned pointers.
__global__ void __launch_bounds__(256) bias_act_vec4_kernel(const float4* __restrict__ x, const float* __restrict__ b,
                                                            const float4* __restrict__ ref, float4* __restrict__ y,
                                                            long long n4, int step_b4, int size_b, int code,
                                                            float alpha, float scale) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    float4 v = x[i];
    if (size_b) {
      const float bb = __ldg(b + (i / step_b4) % size_b);
      v.x = __fadd_rn(v.x, bb); v.y = __fadd_rn(v.y, bb); v.z = __fadd_rn(v.z, bb); v.w = __fadd_rn(v.w, bb);
    }
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ref) r = ref[i];
    float4 o;
    o.x = bias_act_one(v.x, r.x, code, alpha, scale);
    o.y = bias_act_one(v.y, r.y, code, alpha, scale);
    o.z = bias_act_one(v.z, r.z, code, alpha, scale);
    o.w = bias_act_one(v.w, r.w, code, alpha, scale);
    y[i] = o;
  }
}

__global__ void __launch_bounds__(256) bias_act_scalar_kernel(const float* __restrict__ x, const float* __restrict__ b,
                                                              const float* __restrict__ ref, float* __restrict__ y,
                                                              long long n, int step_b, int size_b, int code,
                                                              float alpha, float scale) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    float v = x[i];
    if (size_b) v = __fadd_rn(v, __ldg(b + (i / step_b) % size_b));
    const float r = ref ? ref[i] : 0.f;
    y[i] = bias_act_one(v, r, code, alpha, scale);
  }
}

// x,y: [B,C,H,W]; one thread = 4 consecutive pixels of one (b,c) plane when hw % 4 == 0.
template <int VEC>
__global__ void __launch_bounds__(256) noise_bias_act_kernel(const float* __restrict__ x,
                                                             const float* __restrict__ noise,
                                                             const float* __restrict__ noise_weight,
                                                             const float* __restrict__ bias, float* __restrict__ y,
                                                             long long n_vec, int ch, int hw_vec,
                                                             long long noise_bstride_vec, float slope, float scale) {
  const float nw = noise ? __ldg(noise_weight) : 0.f;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n_vec; i += (long long)gridDim.x * 256) {
    const long long plane = i / hw_vec;
    const int pix = (int)(i - plane * hw_vec);
    const int c = (int)(plane % ch);
    const long long b = plane / ch;
    const float bb = bias ? __ldg(bias + c) : 0.f;
    if (VEC == 4) {
      float4 v = reinterpret_cast<const float4*>(x)[i];
      if (noise) {
        const float4 nz = __ldg(reinterpret_cast<const float4*>(noise) + b * noise_bstride_vec + pix);
        v.x = __fadd_rn(v.x, __fmul_rn(nw, nz.x)); v.y = __fadd_rn(v.y, __fmul_rn(nw, nz.y));
        v.z = __fadd_rn(v.z, __fmul_rn(nw, nz.z)); v.w = __fadd_rn(v.w, __fmul_rn(nw, nz.w));
      }
      float4 o;
      o.x = lrelu_scaled(__fadd_rn(v.x, bb), slope, scale);
      o.y = lrelu_scaled(__fadd_rn(v.y, bb), slope, scale);
      o.z = lrelu_scaled(__fadd_rn(v.z, bb), slope, scale);
      o.w = lrelu_scaled(__fadd_rn(v.w, bb), slope, scale);
      reinterpret_cast<float4*>(y)[i] = o;
    } else {
      float v = x[i];
      if (noise) v = __fadd_rn(v, __fmul_rn(nw, __ldg(noise + b * noise_bstride_vec + pix)));
      y[i] = lrelu_scaled(__fadd_rn(v, bb), slope, scale);
    }
  }
}

static inline unsigned grid_for(long long work_items) {
  long long blocks = (work_items + 255) / 256;
  const long long cap = 148LL * 32;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace maua

extern "C" int maua_fused_bias_act_f32(const float* x, const float* b, const float* ref, float* y, long long n,
                                       int step_b, int size_b, int act, int grad, float alpha, float scale,
                                       void* stream) {
  using namespace maua;
  MAUA_CHECK_ARG(n >= 0 && (n == 0 || (x && y)), "fused_bias_act: null pointer");
  MAUA_CHECK_ARG(size_b >= 0 && (size_b == 0 || (b && step_b >= 1)), "fused_bias_act: bad bias spec");
  if (n == 0) return MAUA_OK;
  const int code = act * 10 + grad;
  cudaStream_t st = as_stream(stream);
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) |
                         reinterpret_cast<uintptr_t>(ref)) & 15) == 0;
  if (aligned && (n % 4 == 0) && (size_b == 0 || step_b % 4 == 0)) {
    bias_act_vec4_kernel<<<grid_for(n / 4), 256, 0, st>>>(
        reinterpret_cast<const float4*>(x), b, reinterpret_cast<const float4*>(ref), reinterpret_cast<float4*>(y),
        n / 4, size_b ? step_b / 4 : 1, size_b, code, alpha, scale);
  } else {
    bias_act_scalar_kernel<<<grid_for(n), 256, 0, st>>>(x, b, ref, y, n, size_b ? step_b : 1, size_b, code, alpha,
                                                         scale);
  }
  MAUA_CHECK_LAUNCH("fused_bias_act");
  return MAUA_OK;
}

extern "C" int maua_noise_bias_act_f32(const float* x, const float* noise, const float* noise_weight,
                                       const float* bias, float* y, int batch, int ch, int h, int w,
                                       long long noise_bstride, float slope, float scale, void* stream) {
  using namespace maua;
  MAUA_CHECK_ARG(x && y && batch >= 0 && ch >= 1 && h >= 1 && w >= 1, "noise_bias_act: bad arguments");
  MAUA_CHECK_ARG(!noise || noise_weight, "noise_bias_act: noise given without noise_weight");
  const long long hw = (long long)h * w;
  const long long n = (long long)batch * ch * hw;
  if (n == 0) return MAUA_OK;
  cudaStream_t st = as_stream(stream);
  const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) |
                         reinterpret_cast<uintptr_t>(noise)) & 15) == 0;
  if (aligned && hw % 4 == 0 && noise_bstride % 4 == 0) {
    noise_bias_act_kernel<4><<<grid_for(n / 4), 256, 0, st>>>(x, noise, noise_weight, bias, y, n / 4, ch,
                                                               (int)(hw / 4), noise_bstride / 4, slope, scale);
  } else {
    noise_bias_act_kernel<1><<<grid_for(n), 256, 0, st>>>(x, noise, noise_weight, bias, y, n, ch, (int)hw,
                                                           noise_bstride, slope, scale);
  }
  MAUA_CHECK_LAUNCH("noise_bias_act");
  return MAUA_OK;
}
