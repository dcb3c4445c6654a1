// upfirdn2d for sm_100a: zero-insert upsample, pad/crop, FIR, decimate — HBM-bound.
//
// Replaces the reference kernels op/upfirdn2d_kernel.cu:49-207 behind the ABI of op/upfirdn2d.cpp:12-23.
// Index semantics (bit-level spec): for output o along one axis
//     mid = o*down + up - 1 - pad0;  in0 = floor(mid/up);  k0 = (in0+1)*up - mid - 1
//     out = sum_y sum_x in[in0y+y, in0x+x] * K[kh-1-(k0y+y*up), kw-1-(k0x+x*up)]      (y-major, x-minor, fp32 FMA)
// Every code path below keeps that accumulation order, so results are bit-identical to the reference CUDA op.
//
// Three kernels:
//   blur_tile_pipe_kernel — the default for identity-rate <= 4x4 FIRs on planes wider than 64: persistent CTAs, the global
//                        loads of tile i+1 in flight while tile i is computed (see the kernel's comment)
//   blur_tile_kernel   — identity-rate (up=down=1), taps <= 4x4, minor == 1 (Blur after the up-conv: 96 % of the
//                        upfirdn2d bytes of a frame).  One CTA = one output tile of one plane; the input tile
//                        (+3 halo) is staged in shared memory with coalesced loads, every thread produces a
//                        4 x RY register block from LDS.128 rows and emits float4 stores.
//   generic_kernel     — everything else (polyphase up/down, large kernels, minor > 1): one thread per output.
#include <cstdlib>

#include "common.cuh"

namespace maua {

struct UfdParams {
  int major, in_h, in_w, minor, kh, kw;
  int up_x, up_y, down_x, down_y, px0, py0;
  int out_h, out_w;
};

// ---------------------------------------------------------------------------------------------------------------
// identity-rate tiled kernel
// ---------------------------------------------------------------------------------------------------------------
template <int TW, int RY, int MINB>
__global__ void __launch_bounds__(256, MINB) blur_tile_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                        const float* __restrict__ k, UfdParams p, int vec_store) {
  constexpr int TXN = TW / 4;        // threads along x
  constexpr int TYN = 256 / TXN;     // thread rows
  constexpr int TH = TYN * RY;       // output rows per tile
  constexpr int COLS = TW + 4;       // staged input columns (TW + 3 needed)
  constexpr int PITCH = TW + 8;      // keeps every row 16-byte aligned
  constexpr int ROWS = TH + 3;
  __shared__ __align__(16) float sx[ROWS * PITCH];
  __shared__ float skf[16];  // flipped, zero-padded to 4x4

  const int tid = threadIdx.x;
  const int tiles_x = (p.out_w + TW - 1) / TW;
  const int tile_x = blockIdx.x % tiles_x;
  const int tile_y = blockIdx.x / tiles_x;
  const int ox0 = tile_x * TW;
  const int oy0 = tile_y * TH;
  const int ix0 = ox0 - p.px0;  // up == 1: in0 = o - pad0, k0 = 0
  const int iy0 = oy0 - p.py0;

  if (tid < 16) {
    const int ky = tid >> 2, kx = tid & 3;
    float v = 0.f;
    if (ky < p.kh && kx < p.kw) v = k[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)];
    skf[tid] = v;
  }

  for (long long plane = blockIdx.y; plane < p.major; plane += gridDim.y) {
    const float* xp = x + plane * (long long)p.in_h * p.in_w;
    __syncthreads();  // previous iteration's readers are done (also orders skf)
    {
      // row-per-warp staging: no integer division, one bounds test per row, coalesced 128-byte warp loads; every
      // thread issues ALL its loads (up to RPW x CPL) before the first shared-memory store, so ~20 independent
      // requests per thread are in flight instead of ~5 (the kernel was latency-bound at 56 % of HBM peak)
      constexpr int RPW = (ROWS + 7) / 8;    // rows per warp
      constexpr int CPL = (COLS + 31) / 32;  // columns per lane
      const int lane = tid & 31, wrp = tid >> 5;
      float v[RPW][CPL];
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        const int r = wrp + 8 * i;
        const int iy = iy0 + r;
        const bool row_ok = (r < ROWS) && (iy >= 0) && (iy < p.in_h);
        const float* src = xp + (long long)iy * p.in_w + ix0;
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const int c = lane + 32 * j;
          const int ix = ix0 + c;
          v[i][j] = (row_ok && c < COLS && ix >= 0 && ix < p.in_w) ? __ldg(src + c) : 0.f;
        }
      }
#pragma unroll
      for (int i = 0; i < RPW; ++i) {
        const int r = wrp + 8 * i;
        if (r < ROWS) {
#pragma unroll
          for (int j = 0; j < CPL; ++j) {
            const int c = lane + 32 * j;
            if (c < COLS) sx[r * PITCH + c] = v[i][j];
          }
        }
      }
    }
    __syncthreads();

    const int tx = tid % TXN;
    const int ty = tid / TXN;
    float kf[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) kf[i] = skf[i];
    float acc[RY][4];
#pragma unroll
    for (int j = 0; j < RY; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;

#pragma unroll
    for (int r = 0; r < RY + 3; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(&sx[(ty * RY + r) * PITCH + tx * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&sx[(ty * RY + r) * PITCH + tx * 4 + 4]);
      const float row[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int j = 0; j < RY; ++j) {
        const int ky = r - j;
        if (ky >= 0 && ky < 4) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) acc[j][i] = __fmaf_rn(row[i + kx], kf[ky * 4 + kx], acc[j][i]);
        }
      }
    }

    float* yp = y + plane * (long long)p.out_h * p.out_w;
    const int ox = ox0 + tx * 4;
#pragma unroll
    for (int j = 0; j < RY; ++j) {
      const int oy = oy0 + ty * RY + j;
      if (oy >= p.out_h) continue;
      float* dst = yp + (long long)oy * p.out_w + ox;
      if (vec_store && ox + 3 < p.out_w) {
        *reinterpret_cast<float4*>(dst) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (ox + i < p.out_w) dst[i] = acc[j][i];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Software-pipelined persistent variant of blur_tile_kernel: a CTA walks tiles with stride gridDim.x and keeps the
// global loads of tile i+1 IN FLIGHT (registers) while it computes tile i out of shared memory.  The one-shot kernel
// alternates load / compute / store phases per CTA, so only a fraction of the resident CTAs have requests outstanding
// at any time (ncu: DRAM 55 %, 0.69 of HBM peak, latency-bound); here every CTA always has ~18 KB outstanding.
// TMA cannot stage these planes: their row pitch (2H+1 floats) is not a multiple of 16 bytes.
// Same tile shape, same FMA order (y-major, x-minor) -> bit-identical results.
// ---------------------------------------------------------------------------------------------------------------
template <int TW, int RY, int MINB>
__global__ void __launch_bounds__(256, MINB) blur_tile_pipe_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                             const float* __restrict__ k, UfdParams p, int vec_store,
                                                             long long n_tiles) {
  constexpr int TXN = TW / 4;
  constexpr int TYN = 256 / TXN;
  constexpr int TH = TYN * RY;
  constexpr int COLS = TW + 4;
  constexpr int PITCH = TW + 8;
  constexpr int ROWS = TH + 3;
  constexpr int RPW = (ROWS + 7) / 8;
  constexpr int CPL = (COLS + 31) / 32;
  __shared__ __align__(16) float sx[ROWS * PITCH];
  __shared__ float skf[16];

  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const int tiles_x = (p.out_w + TW - 1) / TW;
  const int tiles_y = (p.out_h + TH - 1) / TH;
  const long long per_plane = (long long)tiles_x * tiles_y;
  if (tid < 16) {
    const int ky = tid >> 2, kx = tid & 3;
    float v = 0.f;
    if (ky < p.kh && kx < p.kw) v = k[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)];
    skf[tid] = v;
  }
  __syncthreads();
  float kf[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) kf[i] = skf[i];

  float v[RPW][CPL];
  auto load_tile = [&](long long t) {
    const long long plane = t / per_plane;
    const int rem = (int)(t - plane * per_plane);
    const int ix0 = (rem % tiles_x) * TW - p.px0, iy0 = (rem / tiles_x) * TH - p.py0;
    const float* xp = x + plane * (long long)p.in_h * p.in_w;
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
      const int r = wrp + 8 * i;
      const int iy = iy0 + r;
      const bool row_ok = (r < ROWS) && (iy >= 0) && (iy < p.in_h);
      const float* src = xp + (long long)iy * p.in_w + ix0;
#pragma unroll
      for (int j = 0; j < CPL; ++j) {
        const int c = lane + 32 * j;
        const int ix = ix0 + c;
        v[i][j] = (row_ok && c < COLS && ix >= 0 && ix < p.in_w) ? __ldg(src + c) : 0.f;
      }
    }
  };

  long long t = blockIdx.x;
  if (t < n_tiles) load_tile(t);
  const int tx = tid % TXN, ty = tid / TXN;
  for (; t < n_tiles; t += gridDim.x) {
    __syncthreads();  // the previous tile's readers are done with sx
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
      const int r = wrp + 8 * i;
      if (r < ROWS) {
#pragma unroll
        for (int j = 0; j < CPL; ++j) {
          const int c = lane + 32 * j;
          if (c < COLS) sx[r * PITCH + c] = v[i][j];
        }
      }
    }
    __syncthreads();
    const long long tn = t + gridDim.x;
    if (tn < n_tiles) load_tile(tn);   // in flight during the FMA / store phase below

    float acc[RY][4];
#pragma unroll
    for (int j = 0; j < RY; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
#pragma unroll
    for (int r = 0; r < RY + 3; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(&sx[(ty * RY + r) * PITCH + tx * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&sx[(ty * RY + r) * PITCH + tx * 4 + 4]);
      const float row[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int j = 0; j < RY; ++j) {
        const int ky = r - j;
        if (ky >= 0 && ky < 4) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) acc[j][i] = __fmaf_rn(row[i + kx], kf[ky * 4 + kx], acc[j][i]);
        }
      }
    }
    const long long plane = t / per_plane;
    const int rem = (int)(t - plane * per_plane);
    const int ox = (rem % tiles_x) * TW + tx * 4, oy0 = (rem / tiles_x) * TH;
    float* yp = y + plane * (long long)p.out_h * p.out_w;
#pragma unroll
    for (int j = 0; j < RY; ++j) {
      const int oy = oy0 + ty * RY + j;
      if (oy >= p.out_h) continue;
      float* dst = yp + (long long)oy * p.out_w + ox;
      if (vec_store && ox + 3 < p.out_w) {
        *reinterpret_cast<float4*>(dst) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (ox + i < p.out_w) dst[i] = acc[j][i];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// generic kernel: one thread per output element (ox fastest, then minor? no: minor fastest as in memory)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) generic_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                      const float* __restrict__ k, UfdParams p, long long total) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    long long t = i;
    const int m = (int)(t % p.minor);
    t /= p.minor;
    const int ox = (int)(t % p.out_w);
    t /= p.out_w;
    const int oy = (int)(t % p.out_h);
    const long long major = t / p.out_h;

    const int mid_y = oy * p.down_y + p.up_y - 1 - p.py0;
    const int mid_x = ox * p.down_x + p.up_x - 1 - p.px0;
    const int in0_y = floor_div(mid_y, p.up_y);
    const int in0_x = floor_div(mid_x, p.up_x);
    const int k0_y = (in0_y + 1) * p.up_y - mid_y - 1;
    const int k0_x = (in0_x + 1) * p.up_x - mid_x - 1;
    const float* xp = x + major * (long long)p.in_h * p.in_w * p.minor + m;
    float v = 0.f;
    for (int fy = k0_y, iy = in0_y; fy < p.kh; fy += p.up_y, ++iy) {
      if (iy < 0 || iy >= p.in_h) continue;
      const float* krow = k + (p.kh - 1 - fy) * p.kw;
      const float* xrow = xp + (long long)iy * p.in_w * p.minor;
      for (int fx = k0_x, ix = in0_x; fx < p.kw; fx += p.up_x, ++ix) {
        if (ix < 0 || ix >= p.in_w) continue;
        v = __fmaf_rn(__ldg(xrow + (long long)ix * p.minor), __ldg(krow + (p.kw - 1 - fx)), v);
      }
    }
    y[i] = v;
  }
}

}  // namespace maua

extern "C" int maua_upfirdn2d_f32(const float* x, float* y, const float* k, int major, int in_h, int in_w, int minor,
                                  int kh, int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1,
                                  int pad_y0, int pad_y1, void* stream) {
  using namespace maua;
  MAUA_CHECK_ARG(x && y && k, "upfirdn2d: null pointer");
  MAUA_CHECK_ARG(major >= 0 && in_h >= 1 && in_w >= 1 && minor >= 1 && kh >= 1 && kw >= 1, "upfirdn2d: bad shape");
  MAUA_CHECK_ARG(up_x >= 1 && up_y >= 1 && down_x >= 1 && down_y >= 1, "upfirdn2d: up/down must be >= 1");
  UfdParams p;
  p.major = major; p.in_h = in_h; p.in_w = in_w; p.minor = minor; p.kh = kh; p.kw = kw;
  p.up_x = up_x; p.up_y = up_y; p.down_x = down_x; p.down_y = down_y; p.px0 = pad_x0; p.py0 = pad_y0;
  p.out_h = (in_h * up_y + pad_y0 + pad_y1 - kh + down_y) / down_y;
  p.out_w = (in_w * up_x + pad_x0 + pad_x1 - kw + down_x) / down_x;
  if (major == 0 || p.out_h <= 0 || p.out_w <= 0) return MAUA_OK;
  cudaStream_t st = as_stream(stream);

  if (up_x == 1 && up_y == 1 && down_x == 1 && down_y == 1 && kh <= 4 && kw <= 4 && minor == 1) {
    const int vec = ((reinterpret_cast<uintptr_t>(y) & 15) == 0 && (p.out_w & 3) == 0) ? 1 : 0;
    const int planes_y = major < 32768 ? major : 32768;
    const char* ve = getenv("MAUA_UFD_VARIANT");   // read per call (tests / tuning flip it)
    const int variant = ve ? atoi(ve) : 0;
#define MAUA_BLUR_LAUNCH(TWV, RYV, MINBV)                                                    \
  do {                                                                                       \
    constexpr int TH = (256 / (TWV / 4)) * RYV;                                              \
    dim3 grid(ceil_div(p.out_w, TWV) * ceil_div(p.out_h, TH), planes_y);                     \
    blur_tile_kernel<TWV, RYV, MINBV><<<grid, 256, 0, st>>>(x, y, k, p, vec);                \
  } while (0)
    if (p.out_w > 64 && (variant == 0 || variant >= 3)) {
      // software-pipelined persistent kernel.  Measured on B200, [4,32,2049,2049] (4.30 GB in + out, > L2), fraction of the
      // measured HBM copy peak: one-shot tiles 5 CTAs/SM 0.704 | pipelined 2 CTAs/SM 0.684, 3: 0.722, 4: 0.786 (default)
      constexpr int TH = (256 / (128 / 4)) * 4;
      const long long n_tiles = (long long)ceil_div(p.out_w, 128) * ceil_div(p.out_h, TH) * major;
      const int per_sm = variant == 3 ? 3 : (variant == 5 ? 2 : (variant == 6 ? 5 : (variant == 7 ? 6 : 4)));
      long long g = (long long)device_sm_count() * per_sm;
      if (g > n_tiles) g = n_tiles;
      if (per_sm == 2) blur_tile_pipe_kernel<128, 4, 2><<<(unsigned)g, 256, 0, st>>>(x, y, k, p, vec, n_tiles);
      else if (per_sm == 3) blur_tile_pipe_kernel<128, 4, 3><<<(unsigned)g, 256, 0, st>>>(x, y, k, p, vec, n_tiles);
      else if (per_sm == 5) blur_tile_pipe_kernel<128, 4, 5><<<(unsigned)g, 256, 0, st>>>(x, y, k, p, vec, n_tiles);
      else if (per_sm == 6) blur_tile_pipe_kernel<128, 4, 6><<<(unsigned)g, 256, 0, st>>>(x, y, k, p, vec, n_tiles);
      else blur_tile_pipe_kernel<128, 4, 4><<<(unsigned)g, 256, 0, st>>>(x, y, k, p, vec, n_tiles);
    } else if (p.out_w > 64) {
      // measured on B200, [4,32,2049,2049]: (128,4) tiles at 5 CTAs/SM 4.60 TB/s | 4 CTAs 4.24 | (128,8)x2 3.43 |
      // (64,4)x4 3.92 | (128,2)x6 3.93   (MAUA_UFD_VARIANT keeps the alternatives reachable for re-tuning)
      // one-shot tile kernel (MAUA_UFD_VARIANT=1: 5 CTAs/SM, 2: 6 CTAs/SM), kept for A/B measurements
      if (variant == 2) MAUA_BLUR_LAUNCH(128, 4, 6);
      else MAUA_BLUR_LAUNCH(128, 4, 5);
    } else {
      MAUA_BLUR_LAUNCH(32, 2, 4);
    }
#undef MAUA_BLUR_LAUNCH
    MAUA_CHECK_LAUNCH("upfirdn2d(blur_tile)");
    return MAUA_OK;
  }
  const long long total = (long long)major * p.out_h * p.out_w * minor;
  long long blocks = ceil_div(total, 256LL);
  if (blocks > 148LL * 64) blocks = 148LL * 64;
  generic_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, y, k, p, total);
  MAUA_CHECK_LAUNCH("upfirdn2d(generic)");
  return MAUA_OK;
}
