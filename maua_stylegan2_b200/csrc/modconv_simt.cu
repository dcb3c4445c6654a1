// fp32 SIMT modulated convolution — the exact-arithmetic path of ModulatedConv2d (models/stylegan2.py:217-254).
//
// Computes  y[b,co] = d[b,co] * sum_{ci,ky,kx} (c*W[co,ci,ky,kx]) * (s[b,ci] * x[b,ci,..])   
// i.e. the algebraically identical  d (.) conv(x (.) s, W)  form: weights are shared by the whole batch, the style
// is applied to the activation tile as it is staged into shared memory and the demodulation in the epilogue, so the
// reference's [B,Cout,Cin,k,k] weight tensor never exists.  All products/accumulations are fp32 FMA.
//   UP = false: same-resolution conv, zero padding K/2 (K = 1 or 3)
//   UP = true : stride-2 transposed conv, K = 3, output (2H+1)x(2W+1), gather form
//               u[Y,X] = sum_{ky == Y mod 2, kx == X mod 2} W[ky,kx] * x[(Y-ky)/2, (X-kx)/2]
// This kernel is the generic / validation path; the throughput path is modconv_tc.cu (tcgen05).
#include "common.cuh"

namespace maua {

constexpr int CO_TILE = 16;
constexpr int CI_TILE = 8;

template <int K, bool UP, int PX>
__global__ void __launch_bounds__(256) modconv_simt_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ s, const float* __restrict__ d,
                                                           float* __restrict__ y, int cin, int cout, int h, int wd,
                                                           int out_h, int out_w, float w_scale) {
  constexpr int T = 16 * PX;                       // output tile edge
  constexpr int TS = UP ? 10 : (T + K - 1);        // staged input tile edge
  constexpr int PITCH = TS + 1;
  constexpr int KK = K * K;
  __shared__ float xs[CI_TILE][TS][PITCH];
  __shared__ float ws[CO_TILE][CI_TILE][KK];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int tiles_x = (out_w + T - 1) / T;
  const int oy0 = (blockIdx.x / tiles_x) * T;
  const int ox0 = (blockIdx.x % tiles_x) * T;
  const int co0 = blockIdx.y * CO_TILE;
  const int b = blockIdx.z;
  // origin of the staged input tile
  const int iy_base = UP ? (oy0 / 2 - 1) : (oy0 - K / 2);
  const int ix_base = UP ? (ox0 / 2 - 1) : (ox0 - K / 2);

  float acc[PX * PX][CO_TILE];
#pragma unroll
  for (int p = 0; p < PX * PX; ++p)
#pragma unroll
    for (int c = 0; c < CO_TILE; ++c) acc[p][c] = 0.f;

  const float* xb = x + (long long)b * cin * h * wd;
  for (int ci0 = 0; ci0 < cin; ci0 += CI_TILE) {
    __syncthreads();
    for (int idx = tid; idx < CI_TILE * TS * TS; idx += 256) {
      const int cl = idx / (TS * TS);
      const int rem = idx - cl * (TS * TS);
      const int r = rem / TS, c = rem - r * TS;
      const int ci = ci0 + cl, iy = iy_base + r, ix = ix_base + c;
      float v = 0.f;
      if (ci < cin && iy >= 0 && iy < h && ix >= 0 && ix < wd) {
        v = __ldg(xb + ((long long)ci * h + iy) * wd + ix);
        if (s) v *= __ldg(s + (long long)b * cin + ci);
      }
      xs[cl][r][c] = v;
    }
    for (int idx = tid; idx < CO_TILE * CI_TILE * KK; idx += 256) {
      const int col = idx / (CI_TILE * KK);
      const int rem = idx - col * (CI_TILE * KK);
      const int cl = rem / KK, t = rem - cl * KK;
      const int co = co0 + col, ci = ci0 + cl;
      float v = 0.f;
      if (co < cout && ci < cin) v = __ldg(w + ((long long)co * cin + ci) * KK + t) * w_scale;
      ws[col][cl][t] = v;
    }
    __syncthreads();

#pragma unroll 2
    for (int cl = 0; cl < CI_TILE; ++cl) {
      if (!UP) {
        float win[PX + K - 1][PX + K - 1];
#pragma unroll
        for (int r = 0; r < PX + K - 1; ++r)
#pragma unroll
          for (int c = 0; c < PX + K - 1; ++c) win[r][c] = xs[cl][ty * PX + r][tx * PX + c];
#pragma unroll
        for (int col = 0; col < CO_TILE; ++col) {
#pragma unroll
          for (int ky = 0; ky < K; ++ky)
#pragma unroll
            for (int kx = 0; kx < K; ++kx) {
              const float wv = ws[col][cl][ky * K + kx];
#pragma unroll
              for (int j = 0; j < PX; ++j)
#pragma unroll
                for (int i = 0; i < PX; ++i) acc[j * PX + i][col] = fmaf(win[j + ky][i + kx], wv, acc[j * PX + i][col]);
            }
        }
      } else {
        const int Y = oy0 + ty, X = ox0 + tx;
        // taps with (Y - ky) even: Y even -> ky in {0,2}; Y odd -> ky = 1
        const int ky0 = Y & 1, kx0 = X & 1;
        const int nky = (Y & 1) ? 1 : 2, nkx = (X & 1) ? 1 : 2;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          if (a >= nky) break;
          const int ky = ky0 + 2 * a;
          const int ly = ((Y - ky) >> 1) - iy_base;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            if (e >= nkx) break;
            const int kx = kx0 + 2 * e;
            const int lx = ((X - kx) >> 1) - ix_base;
            const float xv = xs[cl][ly][lx];
#pragma unroll
            for (int col = 0; col < CO_TILE; ++col) acc[0][col] = fmaf(xv, ws[col][cl][ky * K + kx], acc[0][col]);
          }
        }
      }
    }
  }

#pragma unroll
  for (int col = 0; col < CO_TILE; ++col) {
    const int co = co0 + col;
    if (co >= cout) break;
    const float dv = d ? __ldg(d + (long long)b * cout + co) : 1.f;
    float* yp = y + ((long long)b * cout + co) * out_h * out_w;
#pragma unroll
    for (int j = 0; j < PX; ++j)
#pragma unroll
      for (int i = 0; i < PX; ++i) {
        const int oy = oy0 + ty * PX + j, ox = ox0 + tx * PX + i;
        if (oy < out_h && ox < out_w) yp[(long long)oy * out_w + ox] = acc[j * PX + i][col] * dv;
      }
  }
}

}  // namespace maua

extern "C" int maua_modconv_simt_f32(const float* x, const float* w, const float* s, const float* d, float* y,
                                     int batch, int cin, int cout, int h, int w_, int ksize, int up, float w_scale,
                                     void* stream) {
  using namespace maua;
  MAUA_CHECK_ARG(x && w && y, "modconv_simt: null pointer");
  MAUA_CHECK_ARG(batch >= 0 && cin >= 1 && cout >= 1 && h >= 1 && w_ >= 1, "modconv_simt: bad shape");
  MAUA_CHECK_ARG(ksize == 1 || ksize == 3, "modconv_simt: kernel size must be 1 or 3");
  MAUA_CHECK_ARG(!up || ksize == 3, "modconv_simt: transposed conv needs kernel size 3");
  if (batch == 0) return MAUA_OK;
  MAUA_CHECK_ARG(batch <= 65535, "modconv_simt: batch too large");
  cudaStream_t st = as_stream(stream);
  const int out_h = up ? 2 * h + 1 : h, out_w = up ? 2 * w_ + 1 : w_;
  if (up) {
    dim3 grid(ceil_div(out_w, 16) * ceil_div(out_h, 16), ceil_div(cout, CO_TILE), batch);
    modconv_simt_kernel<3, true, 1><<<grid, 256, 0, st>>>(x, w, s, d, y, cin, cout, h, w_, out_h, out_w, w_scale);
  } else if (out_h > 16 || out_w > 16) {
    dim3 grid(ceil_div(out_w, 32) * ceil_div(out_h, 32), ceil_div(cout, CO_TILE), batch);
    if (ksize == 3)
      modconv_simt_kernel<3, false, 2><<<grid, 256, 0, st>>>(x, w, s, d, y, cin, cout, h, w_, out_h, out_w, w_scale);
    else
      modconv_simt_kernel<1, false, 2><<<grid, 256, 0, st>>>(x, w, s, d, y, cin, cout, h, w_, out_h, out_w, w_scale);
  } else {
    dim3 grid(ceil_div(out_w, 16) * ceil_div(out_h, 16), ceil_div(cout, CO_TILE), batch);
    if (ksize == 3)
      modconv_simt_kernel<3, false, 1><<<grid, 256, 0, st>>>(x, w, s, d, y, cin, cout, h, w_, out_h, out_w, w_scale);
    else
      modconv_simt_kernel<1, false, 1><<<grid, 256, 0, st>>>(x, w, s, d, y, cin, cout, h, w_, out_h, out_w, w_scale);
  }
  MAUA_CHECK_LAUNCH("modconv_simt");
  return MAUA_OK;
}
