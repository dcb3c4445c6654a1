// Layout / precision-split helpers for the tensor-core path.
//
//   pack_weight_bf16x2   : W[cout,cin,k,k] fp32  ->  [k*k][cout][cin] bf16 (hi, lo), value = W * c  (once per weight)
//   modulate_split_nhwc  : x[B,C,H,W] fp32 * s[b,c]  ->  [B,H,W,C] bf16 (hi, lo)   (network input, or after a bend)
//
// "split" = (hi, lo) with hi = bf16(v), lo = bf16(v - hi): hi + lo carries 16 mantissa bits of v, and the conv
// evaluates hi*hi + hi*lo + lo*hi with fp32 accumulation, which restores ~fp32-grade products (rel. err ~2^-16)
// on the bf16 tensor pipe.
#include "common.cuh"

namespace maua {

// F16 = false: bf16 (hi, lo);  F16 = true: fp16 (hi, lo)
template <bool F16>
__global__ void __launch_bounds__(256) pack_weight_kernel(const float* __restrict__ w, uint16_t* __restrict__ hi,
                                                          uint16_t* __restrict__ lo, int cout, int cin, int kk,
                                                          float scale, long long total) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int ci = (int)(i % cin);
    const long long r = i / cin;
    const int co = (int)(r % cout);
    const int t = (int)(r / cout);
    const float v = __ldg(w + ((long long)co * cin + ci) * kk + t) * scale;
    if (F16) {
      __half h, l;
      split_f16(v, h, l);
      hi[i] = __half_as_ushort(h);
      lo[i] = __half_as_ushort(l);
    } else {
      __nv_bfloat16 h, l;
      split_bf16(v, h, l);
      hi[i] = __bfloat16_as_ushort(h);
      lo[i] = __bfloat16_as_ushort(l);
    }
  }
}

// grid (ceil(HW/32), ceil(C/32), B), block (32, 8): 32x32 (pixel, channel) transpose through shared memory.
// lo == nullptr: single fp16 plane (the "f16" activation format); otherwise the bf16 (hi, lo) pair
__global__ void __launch_bounds__(256) modulate_split_kernel(const float* __restrict__ x, long long x_bstride,
                                                             const float* __restrict__ s,
                                                             uint16_t* __restrict__ hi,
                                                             uint16_t* __restrict__ lo, int ch, long long hw) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const float* xb = x + (long long)b * x_bstride;
#pragma unroll
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j;
    const long long p = p0 + threadIdx.x;
    float v = 0.f;
    if (c < ch && p < hw) {
      v = __ldg(xb + (long long)c * hw + p);
      if (s) v *= __ldg(s + (long long)b * ch + c);
    }
    tile[j][threadIdx.x] = v;  // [channel][pixel]
  }
  __syncthreads();
#pragma unroll
  for (int j = threadIdx.y; j < 32; j += 8) {
    const long long p = p0 + j;
    const int c = c0 + threadIdx.x;
    if (c < ch && p < hw) {
      const long long o = ((long long)b * hw + p) * ch + c;
      if (lo == nullptr) {
        hi[o] = __half_as_ushort(f16_sat(tile[threadIdx.x][j]));
      } else {
        __nv_bfloat16 h, l;
        split_bf16(tile[threadIdx.x][j], h, l);
        hi[o] = __bfloat16_as_ushort(h);
        lo[o] = __bfloat16_as_ushort(l);
      }
    }
  }
}

}  // namespace maua

extern "C" int maua_pack_weight_bf16x2(const float* w, void* w_hi, void* w_lo, int cout, int cin, int ksize,
                                       float w_scale, void* stream) {
  using namespace maua;
  MAUA_CHECK_ARG(w && w_hi && w_lo && cout >= 1 && cin >= 1 && ksize >= 1, "pack_weight: bad arguments");
  const long long total = (long long)ksize * ksize * cout * cin;
  long long blocks = ceil_div(total, 256LL);
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  pack_weight_kernel<false><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
      w, reinterpret_cast<uint16_t*>(w_hi), reinterpret_cast<uint16_t*>(w_lo), cout, cin, ksize * ksize, w_scale, total);
  MAUA_CHECK_LAUNCH("pack_weight");
  return MAUA_OK;
}

extern "C" int maua_pack_weight_f16x2(const float* w, void* w_hi, void* w_lo, int cout, int cin, int ksize,
                                      float w_scale, void* stream) {
  using namespace maua;
  MAUA_CHECK_ARG(w && w_hi && w_lo && cout >= 1 && cin >= 1 && ksize >= 1, "pack_weight_f16x2: bad arguments");
  const long long total = (long long)ksize * ksize * cout * cin;
  long long blocks = ceil_div(total, 256LL);
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  pack_weight_kernel<true><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
      w, reinterpret_cast<uint16_t*>(w_hi), reinterpret_cast<uint16_t*>(w_lo), cout, cin, ksize * ksize, w_scale, total);
  MAUA_CHECK_LAUNCH("pack_weight_f16x2");
  return MAUA_OK;
}

extern "C" int maua_modulate_split_nhwc(const float* x, long long x_bstride, const float* s, void* x_hi, void* x_lo,
                                        int batch, int ch, int h, int w, void* stream) {
  using namespace maua;
  MAUA_CHECK_ARG(x && x_hi && x_lo && batch >= 0 && ch >= 1 && h >= 1 && w >= 1, "modulate_split: bad arguments");
  if (batch == 0) return MAUA_OK;
  MAUA_CHECK_ARG(batch <= 65535 && ch <= 65535 * 32, "modulate_split: shape too large");
  const long long hw = (long long)h * w;
  dim3 grid((unsigned)ceil_div(hw, 32LL), ceil_div(ch, 32), batch);
  modulate_split_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(
      x, x_bstride, s, reinterpret_cast<uint16_t*>(x_hi), reinterpret_cast<uint16_t*>(x_lo), ch, hw);
  MAUA_CHECK_LAUNCH("modulate_split");
  return MAUA_OK;
}

extern "C" int maua_modulate_f16_nhwc(const float* x, long long x_bstride, const float* s, void* x_f16, int batch, int ch,
                                      int h, int w, void* stream) {
  using namespace maua;
  MAUA_CHECK_ARG(x && x_f16 && batch >= 0 && ch >= 1 && h >= 1 && w >= 1, "modulate_f16: bad arguments");
  if (batch == 0) return MAUA_OK;
  MAUA_CHECK_ARG(batch <= 65535 && ch <= 65535 * 32, "modulate_f16: shape too large");
  const long long hw = (long long)h * w;
  dim3 grid((unsigned)ceil_div(hw, 32LL), ceil_div(ch, 32), batch);
  modulate_split_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(x, x_bstride, s, reinterpret_cast<uint16_t*>(x_f16),
                                                                     nullptr, ch, hw);
  MAUA_CHECK_LAUNCH("modulate_f16");
  return MAUA_OK;
}
