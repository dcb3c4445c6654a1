// ToRGB (models/stylegan2.py:356-365) and the frame -> bytes conversion (render.py:40-43).  HBM-bound.
//
//   torgb_kernel: rgb[b,r,p] = sum_c (c*Wrgb[r,c]*s[b,c]) * x[b,c,p] + bias[r] + up2(skip)[b,r,p]
//     - the 1x1 modulated conv has no demodulation (models/stylegan2.py:353), so it is a 3-row GEMV per pixel;
//       the per-sample row vectors live in shared memory, x is streamed once with float4 loads;
//     - up2 = upfirdn2d(skip, K, up=2, pad=(2,1)) evaluated in place with the reference's polyphase index math
//       and FMA order (2x2 live taps per output), so no [B,3,H,W] intermediate is written.
//   rgb_to_u8_kernel: uint8 trunc((clamp(x,-1,1)+1)*127.5), NCHW -> NHWC, 3 B/pixel written.
#include "common.cuh"

namespace maua {

__device__ __forceinline__ float skip_up2(const float* __restrict__ sp, int sh, int sw, const float* kf, int oy, int ox) {
  // up=2, pad0=2, 4x4 taps: mid = o - 1, in0 = floor(mid/2), k0 = 2*in0 + 2 - o  
  const int in0y = floor_div(oy - 1, 2), in0x = floor_div(ox - 1, 2);
  const int k0y = 2 * in0y + 2 - oy, k0x = 2 * in0x + 2 - ox;
  float v = 0.f;
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int iy = in0y + a;
    if (iy < 0 || iy >= sh) continue;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int ix = in0x + e;
      if (ix < 0 || ix >= sw) continue;
      v = __fmaf_rn(__ldg(sp + (long long)iy * sw + ix), kf[(3 - (k0y + 2 * a)) * 4 + (3 - (k0x + 2 * e))], v);
    }
  }
  return v;
}

// skip_up2 for 4 consecutive output pixels ox0 .. ox0+3 (ox0 % 4 == 0) of one row: the 2 skip rows x 4 skip columns
// they touch are loaded once (8 loads instead of 16, no per-pixel index arithmetic); out-of-range taps contribute an
// exact fma(0, k, v) = v, and the FMA order per pixel (row-major over the 2x2 live taps) is skip_up2's, so the results
// are bit-identical.
__device__ __forceinline__ void skip_up2_x4(const float* __restrict__ sp, int sh, int sw, const float* kf, int oy, int ox0,
                                            float out[4]) {
  const int k0y = oy & 1, sy0 = (oy >> 1) - 1 + k0y, sx0 = (ox0 >> 1) - 1;
  float sv[2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int iy = sy0 + a;
    const bool row_ok = iy >= 0 && iy < sh;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int ix = sx0 + c;
      sv[a][c] = (row_ok && ix >= 0 && ix < sw) ? __ldg(sp + (long long)iy * sw + ix) : 0.f;
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k0x = j & 1, cj = (j + 1) >> 1;   // in0x - sx0 = {0,1,1,2}
    float v = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int e = 0; e < 2; ++e) v = __fmaf_rn(sv[a][cj + e], kf[(3 - (k0y + 2 * a)) * 4 + (3 - (k0x + 2 * e))], v);
    out[j] = v;
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) torgb_kernel(const float* __restrict__ x, const float* __restrict__ wrgb,
                                                    const float* __restrict__ s, const float* __restrict__ bias,
                                                    const float* __restrict__ skip, const float* __restrict__ k4,
                                                    float* __restrict__ y, int cin, int h, int w, float w_scale) {
  extern __shared__ float wr[];  // [3][cin]
  __shared__ float kf[16];
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  for (int i = tid; i < 3 * cin; i += 256) {
    const int c = i % cin;
    float v = __ldg(wrgb + i) * w_scale;
    if (s) v *= __ldg(s + (long long)b * cin + c);
    wr[i] = v;
  }
  if (tid < 16) kf[tid] = k4 ? __ldg(k4 + tid) : 0.f;
  __syncthreads();
  const long long hw = (long long)h * w;
  const long long p0 = (blockIdx.x * 256LL + tid) * VEC;
  if (p0 >= hw) return;
  const float* xb = x + (long long)b * cin * hw + p0;
  float acc[3][VEC];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[r][i] = 0.f;
#pragma unroll 4
  for (int c = 0; c < cin; ++c) {
    float xv[VEC];
    if (VEC == 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(xb + (long long)c * hw));
      xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
    } else {
      xv[0] = __ldg(xb + (long long)c * hw);
    }
    const float w0 = wr[c], w1 = wr[cin + c], w2 = wr[2 * cin + c];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      acc[0][i] = fmaf(xv[i], w0, acc[0][i]);
      acc[1][i] = fmaf(xv[i], w1, acc[1][i]);
      acc[2][i] = fmaf(xv[i], w2, acc[2][i]);
    }
  }
  const int sh = h >> 1, sw = w >> 1;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const float bb = bias ? __ldg(bias + r) : 0.f;
    float o[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      o[i] = __fadd_rn(acc[r][i], bb);
      if (skip) {
        const long long p = p0 + i;
        const int oy = (int)(p / w), ox = (int)(p - (long long)oy * w);
        o[i] = __fadd_rn(o[i], skip_up2(skip + ((long long)b * 3 + r) * sh * sw, sh, sw, kf, oy, ox));
      }
    }
    float* yp = y + ((long long)b * 3 + r) * hw + p0;
    if (VEC == 4) *reinterpret_cast<float4*>(yp) = make_float4(o[0], o[1], o[2], o[3]);
    else yp[0] = o[0];
  }
}

// Low-resolution variant (H*W <= 4096, Cin = 512): the serial channel loop of torgb_kernel is latency-bound there
// (a handful of threads, 512 dependent iterations), so the channels are split over the 8 warps of the block
// (32 pixels x 8 channel groups) and the partial sums are combined through shared memory.
__global__ void __launch_bounds__(256) torgb_small_kernel(const float* __restrict__ x, const float* __restrict__ wrgb,
                                                          const float* __restrict__ s, const float* __restrict__ bias,
                                                          const float* __restrict__ skip, const float* __restrict__ k4,
                                                          float* __restrict__ y, int cin, int h, int w, float w_scale) {
  extern __shared__ float wr[];  // [3][cin]
  __shared__ float kf[16];
  __shared__ float part[3][8][32];
  const int b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, cg = tid >> 5;
  for (int i = tid; i < 3 * cin; i += 256) {
    const int c = i % cin;
    float v = __ldg(wrgb + i) * w_scale;
    if (s) v *= __ldg(s + (long long)b * cin + c);
    wr[i] = v;
  }
  if (tid < 16) kf[tid] = k4 ? __ldg(k4 + tid) : 0.f;
  __syncthreads();
  const int hw = h * w;
  const int p = blockIdx.x * 32 + lane;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  if (p < hw) {
    const float* xb = x + (long long)b * cin * hw + p;
    // same accumulation order per partial (ascending c), partials combined in ascending group order below.
    // Eight activation loads are issued before the first dependent FMA: with one load per iteration the 64-step chain
    // paid the L2 latency 64 times (ncu: 17 us for the 8x8 layer, whose data is 1 MB).
    for (int c0 = cg; c0 < cin; c0 += 64) {
      float xv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int c = c0 + 8 * u;
        xv[u] = c < cin ? __ldg(xb + (long long)c * hw) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int c = c0 + 8 * u;
        if (c < cin) {
          a0 = fmaf(xv[u], wr[c], a0);
          a1 = fmaf(xv[u], wr[cin + c], a1);
          a2 = fmaf(xv[u], wr[2 * cin + c], a2);
        }
      }
    }
  }
  part[0][cg][lane] = a0;
  part[1][cg][lane] = a1;
  part[2][cg][lane] = a2;
  __syncthreads();
  if (cg < 3 && p < hw) {
    const int r = cg;
    float acc = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) acc += part[r][g][lane];
    float o = __fadd_rn(acc, bias ? __ldg(bias + r) : 0.f);
    if (skip) {
      const int sh = h >> 1, sw = w >> 1;
      const int oy = p / w, ox = p - oy * w;
      o = __fadd_rn(o, skip_up2(skip + ((long long)b * 3 + r) * sh * sw, sh, sw, kf, oy, ox));
    }
    y[((long long)b * 3 + r) * hw + p] = o;
  }
}

__global__ void __launch_bounds__(256) rgb_weights_kernel(const float* __restrict__ wrgb, const float* __restrict__ s,
                                                          float* __restrict__ wr, int cin, float w_scale, int total) {
  const int i = blockIdx.x * 256 + threadIdx.x;  // over [B][3][cin]
  if (i >= total) return;
  const int c = i % cin, k = (i / cin) % 3, b = i / (3 * cin);
  float v = __ldg(wrgb + k * cin + c) * w_scale;
  if (s) v *= __ldg(s + (long long)b * cin + c);
  wr[i] = v;
}

// VEC = 4: one thread = 4 consecutive pixels of one (b, k) plane (float4 in / out); the 2x2 live taps of the up-2
// FIR come from 3 skip columns x 2 skip rows held in registers.  Rounding identical to skip_up2 / torgb_kernel.
template <int VEC>
__global__ void __launch_bounds__(256) rgb_finish_kernel(const float* __restrict__ partial, const float* __restrict__ bias,
                                                         const float* __restrict__ skip, const float* __restrict__ k4,
                                                         float* __restrict__ y, int h, int w, long long total_vec) {
  __shared__ float kf[16];
  if (threadIdx.x < 16) kf[threadIdx.x] = k4 ? __ldg(k4 + threadIdx.x) : 0.f;
  __syncthreads();
  const int wv = w / VEC;
  const long long hwv = (long long)h * wv;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total_vec; i += (long long)gridDim.x * 256) {
    const long long plane = i / hwv;  // b*3 + k
    const int rem = (int)(i - plane * hwv);
    const int oy = rem / wv, ox0 = (rem - oy * wv) * VEC;
    const float bb = bias ? __ldg(bias + (int)(plane % 3)) : 0.f;
    float o[VEC];
    if (VEC == 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(partial) + i);
      o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
    } else {
      o[0] = partial[i];
    }
    const float* sp = skip ? skip + plane * (h >> 1) * (w >> 1) : nullptr;
    float sk[4] = {0.f, 0.f, 0.f, 0.f};
    if (skip && VEC == 4) skip_up2_x4(sp, h >> 1, w >> 1, kf, oy, ox0, sk);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      o[j] = __fadd_rn(o[j], bb);
      if (skip) o[j] = __fadd_rn(o[j], VEC == 4 ? sk[j] : skip_up2(sp, h >> 1, w >> 1, kf, oy, ox0 + j));
    }
    if (VEC == 4) reinterpret_cast<float4*>(y)[i] = make_float4(o[0], o[1], o[2], o[3]);
    else y[i] = o[0];
  }
}

// Last ToRGB of a frame straight to bytes: (partial + bias) + up2(skip) for the three planes of 4 consecutive pixels, then
// render.py:40-43's clamp / scale / truncate and the NCHW -> NHWC interleave — 12 bytes out per thread.  The fp32 image of
// the final resolution (100 MB per batch of 8 at 1024^2) is never written or re-read.  Same rounding as
// rgb_finish_kernel followed by rgb_to_u8_kernel.
__global__ void __launch_bounds__(256) rgb_finish_u8_kernel(const float* __restrict__ partial, const float* __restrict__ bias,
                                                            const float* __restrict__ skip, const float* __restrict__ k4,
                                                            uint8_t* __restrict__ out, int h, int w, long long total_vec) {
  __shared__ float kf[16];
  if (threadIdx.x < 16) kf[threadIdx.x] = k4 ? __ldg(k4 + threadIdx.x) : 0.f;
  __syncthreads();
  const int wv = w / 4;
  const long long hwv = (long long)h * wv, hw = (long long)h * w;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total_vec; i += (long long)gridDim.x * 256) {
    const long long b = i / hwv;
    const int rem = (int)(i - b * hwv);
    const int oy = rem / wv, ox0 = (rem - oy * wv) * 4;
    uint8_t px[12];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(partial + (b * 3 + k) * hw + (long long)oy * w + ox0));
      float o[4] = {t.x, t.y, t.z, t.w};
      const float bb = bias ? __ldg(bias + k) : 0.f;
      const float* sp = skip ? skip + (b * 3 + k) * (h >> 1) * (w >> 1) : nullptr;
      float sk[4] = {0.f, 0.f, 0.f, 0.f};
      if (skip) skip_up2_x4(sp, h >> 1, w >> 1, kf, oy, ox0, sk);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = __fadd_rn(o[j], bb);
        if (skip) v = __fadd_rn(v, sk[j]);
        v = fminf(fmaxf(v, -1.f), 1.f);
        v = __fmul_rn(__fadd_rn(v, 1.f), 127.5f);
        px[j * 3 + k] = (uint8_t)(int)v;
      }
    }
    uint32_t* dst = reinterpret_cast<uint32_t*>(out + ((b * h + oy) * (long long)w + ox0) * 3);   // 12-byte aligned: ox0 % 4 == 0
    dst[0] = px[0] | (px[1] << 8) | (px[2] << 16) | ((uint32_t)px[3] << 24);
    dst[1] = px[4] | (px[5] << 8) | (px[6] << 16) | ((uint32_t)px[7] << 24);
    dst[2] = px[8] | (px[9] << 8) | (px[10] << 16) | ((uint32_t)px[11] << 24);
  }
}

// one thread = one pixel (3 planar loads coalesced across the warp, 3 packed bytes out)
__global__ void __launch_bounds__(256) rgb_to_u8_kernel(const float* __restrict__ rgb, uint8_t* __restrict__ out,
                                                        long long hw, long long total) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long b = i / hw, p = i - b * hw;
    const float* src = rgb + b * 3 * hw + p;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = __ldg(src + c * hw);
      v = fminf(fmaxf(v, -1.f), 1.f);
      v = __fmul_rn(__fadd_rn(v, 1.f), 127.5f);
      out[i * 3 + c] = (uint8_t)(int)v;  // truncation toward zero == numpy astype(uint8) on [0,255]
    }
  }
}

// Crop + two-pass fixed-point bilinear resample of uint8 NHWC frames, bit-for-bit what the reference's frame loop does
// on the host per frame with PIL (render.py:98-105: img[:, 112:-112] -> Image.resize((1920, 1080), BILINEAR)):
// Pillow's ImagingResample = horizontal pass then vertical pass, 22-bit integer coefficients, each pass rounded to
// 8 bits ((1 << 21) + sum) >> 22, clipped).  One thread = one output pixel: it re-derives the <= ksize_y horizontally
// resampled values it needs (upscaling: 3x3 source bytes per channel), so no intermediate image touches HBM.
__global__ void __launch_bounds__(256) fit_frames_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                         int in_h, int in_w, int cy0, int cx0, int oh, int ow,
                                                         const int* __restrict__ bx, const int* __restrict__ kx, int ksx,
                                                         const int* __restrict__ by, const int* __restrict__ ky, int ksy,
                                                         long long total) {
  constexpr int PB = 22;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int xx = (int)(i % ow);
    const long long r = i / ow;
    const int yy = (int)(r % oh);
    const long long b = r / oh;
    const int x0 = __ldg(bx + 2 * xx) + cx0, nx = __ldg(bx + 2 * xx + 1);
    const int y0 = __ldg(by + 2 * yy) + cy0, ny = __ldg(by + 2 * yy + 1);
    const uint8_t* src = in + b * (long long)in_h * in_w * 3;
    int acc[3] = {1 << (PB - 1), 1 << (PB - 1), 1 << (PB - 1)};
    for (int j = 0; j < ny; ++j) {
      const uint8_t* row = src + ((long long)(y0 + j) * in_w + x0) * 3;
      int h[3] = {1 << (PB - 1), 1 << (PB - 1), 1 << (PB - 1)};
      for (int t = 0; t < nx; ++t) {
        const int k = __ldg(kx + xx * ksx + t);
#pragma unroll
        for (int c = 0; c < 3; ++c) h[c] += (int)__ldg(row + t * 3 + c) * k;
      }
      const int kv = __ldg(ky + yy * ksy + j);
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[c] += min(max(h[c] >> PB, 0), 255) * kv;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) out[i * 3 + c] = (uint8_t)min(max(acc[c] >> PB, 0), 255);
  }
}

}  // namespace maua

extern "C" int maua_fit_frames_u8(const uint8_t* in, uint8_t* out, int batch, int in_h, int in_w, int crop_y0,
                                  int crop_x0, int crop_h, int crop_w, int out_h, int out_w, const int* bounds_x,
                                  const int* coef_x, int ksize_x, const int* bounds_y, const int* coef_y, int ksize_y,
                                  void* stream) {
  using namespace maua;
  MAUA_CHECK_ARG(in && out && in != out && bounds_x && coef_x && bounds_y && coef_y, "fit_frames: null or aliased pointers");
  MAUA_CHECK_ARG(batch >= 0 && in_h >= 1 && in_w >= 1 && out_h >= 1 && out_w >= 1 && ksize_x >= 1 && ksize_y >= 1,
                 "fit_frames: bad shape");
  MAUA_CHECK_ARG(crop_y0 >= 0 && crop_x0 >= 0 && crop_h >= 1 && crop_w >= 1 && crop_y0 + crop_h <= in_h &&
                     crop_x0 + crop_w <= in_w,
                 "fit_frames: crop window outside the frame");
  const long long total = (long long)batch * out_h * out_w;
  if (total == 0) return MAUA_OK;
  long long blocks = ceil_div(total, 256LL);
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  fit_frames_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(in, out, in_h, in_w, crop_y0, crop_x0, out_h, out_w,
                                                                    bounds_x, coef_x, ksize_x, bounds_y, coef_y,
                                                                    ksize_y, total);
  MAUA_CHECK_LAUNCH("fit_frames");
  return MAUA_OK;
}

extern "C" int maua_torgb_f32(const float* x, const float* wrgb, const float* s, const float* bias,
                              const float* skip, const float* k4, float* y, int batch, int cin, int h, int w,
                              float w_scale, void* stream) {
  using namespace maua;
  MAUA_CHECK_ARG(x && wrgb && y && batch >= 0 && cin >= 1 && h >= 1 && w >= 1, "torgb: bad arguments");
  MAUA_CHECK_ARG(!skip || (k4 && (h % 2 == 0) && (w % 2 == 0)), "torgb: skip needs k4 and even output size");
  MAUA_CHECK_ARG(cin <= 4096, "torgb: cin too large");
  if (batch == 0) return MAUA_OK;
  MAUA_CHECK_ARG(batch <= 65535, "torgb: batch too large");
  const long long hw = (long long)h * w;
  cudaStream_t st = as_stream(stream);
  const size_t smem = 3 * (size_t)cin * sizeof(float);
  const bool vec = (hw % 4 == 0) && (w % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  if (hw <= 16384 && cin >= 64) {  // channel-split kernel: the serial-channel kernel ran the 256-ch 128^2 ToRGB at 22 % DRAM
    dim3 grid((unsigned)ceil_div(hw, 32LL), batch);
    torgb_small_kernel<<<grid, 256, smem, st>>>(x, wrgb, s, bias, skip, k4, y, cin, h, w, w_scale);
  } else if (vec) {
    dim3 grid((unsigned)ceil_div(hw / 4, 256LL), batch);
    torgb_kernel<4><<<grid, 256, smem, st>>>(x, wrgb, s, bias, skip, k4, y, cin, h, w, w_scale);
  } else {
    dim3 grid((unsigned)ceil_div(hw, 256LL), batch);
    torgb_kernel<1><<<grid, 256, smem, st>>>(x, wrgb, s, bias, skip, k4, y, cin, h, w, w_scale);
  }
  MAUA_CHECK_LAUNCH("torgb");
  return MAUA_OK;
}

extern "C" int maua_rgb_to_u8_nhwc(const float* rgb, uint8_t* out, int batch, int h, int w, void* stream) {
  using namespace maua;
  MAUA_CHECK_ARG(rgb && out && batch >= 0 && h >= 1 && w >= 1, "rgb_to_u8: bad arguments");
  const long long hw = (long long)h * w, total = hw * batch;
  if (total == 0) return MAUA_OK;
  long long blocks = ceil_div(total, 256LL);
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  rgb_to_u8_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(rgb, out, hw, total);
  MAUA_CHECK_LAUNCH("rgb_to_u8");
  return MAUA_OK;
}

extern "C" int maua_rgb_weights_f32(const float* wrgb, const float* s, float* wr, int batch, int cin, float w_scale,
                                    void* stream) {
  using namespace maua;
  MAUA_CHECK_ARG(wrgb && wr && batch >= 0 && cin >= 1, "rgb_weights: bad arguments");
  const int total = batch * 3 * cin;
  if (total == 0) return MAUA_OK;
  rgb_weights_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(wrgb, s, wr, cin, w_scale, total);
  MAUA_CHECK_LAUNCH("rgb_weights");
  return MAUA_OK;
}

extern "C" int maua_rgb_finish_f32(const float* partial, const float* bias, const float* skip, const float* k4, float* y,
                                   int batch, int h, int w, void* stream) {
  using namespace maua;
  MAUA_CHECK_ARG(partial && y && batch >= 0 && h >= 1 && w >= 1, "rgb_finish: bad arguments");
  MAUA_CHECK_ARG(!skip || (k4 && (h % 2 == 0) && (w % 2 == 0)), "rgb_finish: skip needs k4 and even output size");
  const long long total = (long long)batch * 3 * h * w;
  if (total == 0) return MAUA_OK;
  const bool vec = (w % 4 == 0) && ((reinterpret_cast<uintptr_t>(partial) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
  const long long work = vec ? total / 4 : total;
  long long blocks = ceil_div(work, 256LL);
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  if (vec) rgb_finish_kernel<4><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(partial, bias, skip, k4, y, h, w, work);
  else rgb_finish_kernel<1><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(partial, bias, skip, k4, y, h, w, work);
  MAUA_CHECK_LAUNCH("rgb_finish");
  return MAUA_OK;
}

extern "C" int maua_rgb_finish_u8(const float* partial, const float* bias, const float* skip, const float* k4, uint8_t* out,
                                  int batch, int h, int w, void* stream) {
  using namespace maua;
  MAUA_CHECK_ARG(partial && out && batch >= 0 && h >= 1 && w >= 1, "rgb_finish_u8: bad arguments");
  MAUA_CHECK_ARG(!skip || (k4 && (h % 2 == 0) && (w % 2 == 0)), "rgb_finish_u8: skip needs k4 and even output size");
  MAUA_CHECK_ARG(w % 4 == 0 && (reinterpret_cast<uintptr_t>(partial) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0,
                 "rgb_finish_u8: width must be a multiple of 4 and the buffers 16 / 4 byte aligned");
  const long long work = (long long)batch * h * (w / 4);
  if (work == 0) return MAUA_OK;
  long long blocks = ceil_div(work, 256LL);
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  rgb_finish_u8_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(partial, bias, skip, k4, out, h, w, work);
  MAUA_CHECK_LAUNCH("rgb_finish_u8");
  return MAUA_OK;
}
