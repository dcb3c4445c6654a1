// Network-bend warps and looping Perlin noise.
//
//   maua_bend_warp_f32      <- audioreactive/bend.py:51-102  Translate / Zoom / Rotate: the reference builds
//                              nn.Sequential(ReflectionPad2d x1..3, AddNoise, kornia Translate|Scale|Rotate, CenterCrop),
//                              i.e. 4-6 full HBM round trips over a tensor padded to 5x (Translate) or 9x (Zoom) its size.
//                              Here ONE kernel evaluates the chain per OUTPUT pixel: crop offset -> inverse affine ->
//                              bilinear taps (zeros outside the padded frame) -> each tap's padded index is folded back
//                              through the pad stages to an input index (+ the additive noise at the padded index).
//                              HBM traffic = read x once (L2-served re-reads) + write y once.
//   maua_perlin_noise       <- audioreactive/latent.py:188-246 perlin_noise: the reference materialises eight
//                              [T,H,W,3] float64 gradient tensors (repeat + slice) and ~30 temporaries; here the lattice
//                              gradients [r0+1,r1+1,r2+1,3] stay in L1/L2 and every voxel is computed in registers in
//                              the reference's operation order (fp64, no contraction) -> 8 or 4 B/voxel of HBM writes.
#include "common.cuh"

namespace maua {
namespace bend {

struct PadAxis {
  int n;         // number of stages (0..4)
  int left[4];   // left/top pad of stage s
  int size[4];   // extent of the axis BEFORE stage s
};

// padded index -> input index through the pad stages (last stage first); mode 0 reflect, 1 replicate
__device__ __forceinline__ int fold(int j, const PadAxis& p, int mode) {
#pragma unroll
  for (int s = 3; s >= 0; --s) {
    if (s >= p.n) continue;
    j -= p.left[s];
    const int n = p.size[s];
    if (mode == 0) {
      if (j < 0) j = -j;
      if (j >= n) j = 2 * (n - 1) - j;
    } else {
      j = j < 0 ? 0 : (j >= n ? n - 1 : j);
    }
  }
  return j;
}

constexpr int CH_PER_THREAD = 8;

__global__ void __launch_bounds__(256) bend_warp_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                        const float* __restrict__ noise,
                                                        const float* __restrict__ minv, int B, int C, int H, int W,
                                                        PadAxis px, PadAxis py, int mode, int Hp, int Wp, int noise_b,
                                                        int noise_c, int OH, int OW, int cy0, int cx0) {
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * CH_PER_THREAD;
  const int pix = blockIdx.x * 256 + threadIdx.x;
  if (pix >= OH * OW) return;
  const int oy = pix / OW, ox = pix - oy * OW;
  const float* m = minv + b * 6;
  const float X = (float)(ox + cx0), Y = (float)(oy + cy0);
  // source position in the padded frame (kornia warp_affine: dst = M src, sampled at src = M^-1 dst, align_corners)
  const float sx = fmaf(__ldg(m + 0), X, fmaf(__ldg(m + 1), Y, __ldg(m + 2)));
  const float sy = fmaf(__ldg(m + 3), X, fmaf(__ldg(m + 4), Y, __ldg(m + 5)));
  const float fx0 = floorf(sx), fy0 = floorf(sy);
  const int x0 = (int)fx0, y0 = (int)fy0;
  const float ax = sx - fx0, ay = sy - fy0;
  // grid_sample's corner weights (nw, ne, sw, se)
  const float wgt[4] = {(1.f - ax) * (1.f - ay), ax * (1.f - ay), (1.f - ax) * ay, ax * ay};
  int src[4], nz[4];
  bool ok[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int qx = x0 + (t & 1), qy = y0 + (t >> 1);
    ok[t] = qx >= 0 && qx < Wp && qy >= 0 && qy < Hp;
    src[t] = ok[t] ? fold(qy, py, mode) * W + fold(qx, px, mode) : 0;
    nz[t] = ok[t] ? qy * Wp + qx : 0;
  }
  const long long nbs = noise_b > 1 ? (long long)noise_c * Hp * Wp : 0;
#pragma unroll
  for (int k = 0; k < CH_PER_THREAD; ++k) {
    const int c = c0 + k;
    if (c >= C) break;
    const float* xc = x + ((long long)b * C + c) * H * W;
    const float* nc = noise ? noise + b * nbs + (noise_c > 1 ? (long long)c * Hp * Wp : 0) : nullptr;
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (!ok[t]) continue;
      float v = __ldg(xc + src[t]);
      if (nc) v += __ldg(nc + nz[t]);
      acc = fmaf(v, wgt[t], acc);
    }
    y[(((long long)b * C + c) * OH + oy) * OW + ox] = acc;
  }
}

// ---- Perlin -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double dot3(const double* __restrict__ g, double a, double b, double c) {
  // th.sum(stack((a, b, c)) * g, 3): three rounded products, summed left to right (no FMA contraction)
  return __dadd_rn(__dadd_rn(__dmul_rn(a, __ldg(g + 0)), __dmul_rn(b, __ldg(g + 1))), __dmul_rn(c, __ldg(g + 2)));
}
__device__ __forceinline__ double fade(double t) {
  // t * t * t * (t * (t * 6 - 15) + 10), evaluated in Python's order
  const double inner = __dadd_rn(__dmul_rn(t, __dadd_rn(__dmul_rn(t, 6.0), -15.0)), 10.0);
  return __dmul_rn(__dmul_rn(__dmul_rn(t, t), t), inner);
}
__device__ __forceinline__ double lerp_a(double a, double b, double t) {  // a * (1 - t) + t * b
  return __dadd_rn(__dmul_rn(a, __dadd_rn(1.0, -t)), __dmul_rn(t, b));
}

template <typename OutT>
__global__ void __launch_bounds__(256) perlin_kernel(const double* __restrict__ grad, OutT* __restrict__ out, int s0,
                                                     int s1, int s2, int r0, int r1, int r2, double dl0, double dl1,
                                                     double dl2) {
  const long long total = (long long)s0 * s1 * s2;
  const int d0 = s0 / r0, d1 = s1 / r1, d2 = s2 / r2;
  const int g1 = r1 + 1, g2 = r2 + 1;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int k = (int)(i % s2);
    const long long ij = i / s2;
    const int j = (int)(ij % s1), t = (int)(ij / s1);
    // np.mgrid[0:res:delta] % 1  ->  fmod(index * delta, 1)
    const double p0 = fmod(__dmul_rn((double)t, dl0), 1.0), p1 = fmod(__dmul_rn((double)j, dl1), 1.0),
                 p2 = fmod(__dmul_rn((double)k, dl2), 1.0);
    const int c0 = t / d0, c1 = j / d1, c2 = k / d2;  // gradients.repeat(d) + slicing == lattice[index // d (+1)]
    const double* g = grad + (((long long)c0 * g1 + c1) * g2 + c2) * 3;
    const long long st0 = (long long)g1 * g2 * 3, st1 = (long long)g2 * 3, st2 = 3;
    const double n000 = dot3(g, p0, p1, p2);
    const double n100 = dot3(g + st0, p0 - 1.0, p1, p2);
    const double n010 = dot3(g + st1, p0, p1 - 1.0, p2);
    const double n110 = dot3(g + st0 + st1, p0 - 1.0, p1 - 1.0, p2);
    const double n001 = dot3(g + st2, p0, p1, p2 - 1.0);
    const double n101 = dot3(g + st0 + st2, p0 - 1.0, p1, p2 - 1.0);
    const double n011 = dot3(g + st1 + st2, p0, p1 - 1.0, p2 - 1.0);
    const double n111 = dot3(g + st0 + st1 + st2, p0 - 1.0, p1 - 1.0, p2 - 1.0);
    const double t0 = fade(p0), t1 = fade(p1), t2 = fade(p2);
    const double n00 = lerp_a(n000, n100, t0), n10 = lerp_a(n010, n110, t0);
    const double n01 = lerp_a(n001, n101, t0), n11 = lerp_a(n011, n111, t0);
    // (1 - t1) * n00 + t1 * n10
    const double n0 = __dadd_rn(__dmul_rn(__dadd_rn(1.0, -t1), n00), __dmul_rn(t1, n10));
    const double n1 = __dadd_rn(__dmul_rn(__dadd_rn(1.0, -t1), n01), __dmul_rn(t1, n11));
    const double v = __dadd_rn(__dmul_rn(__dadd_rn(1.0, -t2), n0), __dmul_rn(t2, n1));
    out[i] = (OutT)__dadd_rn(__dmul_rn(v, 2.0), -1.0);
  }
}

}  // namespace bend
}  // namespace maua

using namespace maua;
using namespace maua::bend;

static bool make_axis(PadAxis& a, const int* pads, int n, int extent, int mode, int* padded) {
  a.n = n;
  for (int s = 0; s < 4; ++s) a.left[s] = 0, a.size[s] = 1;
  for (int s = 0; s < n; ++s) {
    const int l = pads[2 * s], r = pads[2 * s + 1];
    if (l < 0 || r < 0) return false;
    if (mode == 0 && (l >= extent || r >= extent)) return false;  // torch ReflectionPad2d: pad < input size
    a.left[s] = l;
    a.size[s] = extent;
    extent += l + r;
  }
  *padded = extent;
  return true;
}

extern "C" int maua_bend_warp_f32(const float* x, float* y, const float* noise, const float* minv, int batch, int ch,
                                  int h, int w, const int* pad_x_host, int n_pad_x, const int* pad_y_host, int n_pad_y,
                                  int pad_mode, int noise_b, int noise_c, int out_h, int out_w, int crop_y0,
                                  int crop_x0, void* stream) {
  MAUA_CHECK_ARG(x && y && minv && x != y, "bend_warp: null or aliased pointers");
  MAUA_CHECK_ARG(batch >= 1 && batch <= 65535 && ch >= 1 && h >= 1 && w >= 1 && out_h >= 1 && out_w >= 1,
                 "bend_warp: bad shape");
  MAUA_CHECK_ARG(n_pad_x >= 0 && n_pad_x <= 4 && n_pad_y >= 0 && n_pad_y <= 4, "bend_warp: at most 4 pad stages per axis");
  MAUA_CHECK_ARG((n_pad_x == 0 || pad_x_host) && (n_pad_y == 0 || pad_y_host), "bend_warp: pad arrays missing");
  MAUA_CHECK_ARG(pad_mode == 0 || pad_mode == 1, "bend_warp: pad_mode 0 (reflect) or 1 (replicate)");
  PadAxis px, py;
  int hp = h, wp = w;
  MAUA_CHECK_ARG(make_axis(px, pad_x_host, n_pad_x, w, pad_mode, &wp), "bend_warp: x padding must be >= 0 and (reflect) smaller than the axis");
  MAUA_CHECK_ARG(make_axis(py, pad_y_host, n_pad_y, h, pad_mode, &hp), "bend_warp: y padding must be >= 0 and (reflect) smaller than the axis");
  MAUA_CHECK_ARG(!noise || ((noise_b == 1 || noise_b == batch) && (noise_c == 1 || noise_c == ch)),
                 "bend_warp: noise must broadcast over batch / channels");
  MAUA_CHECK_ARG(crop_y0 >= 0 && crop_x0 >= 0 && crop_y0 + out_h <= hp && crop_x0 + out_w <= wp,
                 "bend_warp: crop window [%d:%d, %d:%d] outside the padded frame %dx%d", crop_y0, crop_y0 + out_h,
                 crop_x0, crop_x0 + out_w, hp, wp);
  dim3 grid(ceil_div(out_h * out_w, 256), ceil_div(ch, CH_PER_THREAD), batch);
  bend_warp_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, y, noise, minv, batch, ch, h, w, px, py, pad_mode, hp, wp,
                                                        noise_b, noise_c, out_h, out_w, crop_y0, crop_x0);
  MAUA_CHECK_LAUNCH("bend_warp");
  return MAUA_OK;
}

extern "C" int maua_perlin_noise(const double* gradients, void* out, int s0, int s1, int s2, int r0, int r1, int r2,
                                 int out_f64, void* stream) {
  MAUA_CHECK_ARG(gradients && out, "perlin_noise: null pointers");
  MAUA_CHECK_ARG(r0 >= 1 && r1 >= 1 && r2 >= 1 && s0 >= r0 && s1 >= r1 && s2 >= r2, "perlin_noise: bad shape/res");
  MAUA_CHECK_ARG(s0 % r0 == 0 && s1 % r1 == 0 && s2 % r2 == 0, "perlin_noise: shape must be a multiple of res");
  const long long total = (long long)s0 * s1 * s2;
  long long blocks = ceil_div(total, 256LL);
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  const double dl0 = (double)r0 / s0, dl1 = (double)r1 / s1, dl2 = (double)r2 / s2;
  if (out_f64)
    perlin_kernel<double><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(gradients, (double*)out, s0, s1, s2, r0, r1,
                                                                           r2, dl0, dl1, dl2);
  else
    perlin_kernel<float><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(gradients, (float*)out, s0, s1, s2, r0, r1, r2,
                                                                          dl0, dl1, dl2);
  MAUA_CHECK_LAUNCH("perlin_noise");
  return MAUA_OK;
}
