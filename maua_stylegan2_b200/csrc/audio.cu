// cuFFT-fronted audio feature chain for sm_100a (replaces the librosa/scipy CPU path of audioreactive/signal.py:31-156
// and the glue of examples/default.py:6-45).  Everything stays on the device: audio -> STFT -> HPSS -> ISTFT -> STFT ->
// mel/onset | chroma/CENS/kNN-median | rms -> Fourier resample -> gaussian filter -> percentile clip -> latents.
// One-shot preprocessing (seconds of audio per millisecond), HBM/cuFFT-bound; kernels favour clarity + coalescing.
//
// Spectrogram layout: S[t][f] interleaved (re, im) fp32, F = n_fft/2 + 1, t-major (one cuFFT batch row per frame).
#include <cufft.h>

#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace maua {
namespace audio {

constexpr float kPi = 3.14159265358979323846f;

// ---------------------------------------------------------------------------------------------------------------
// cuFFT plan cache
// ---------------------------------------------------------------------------------------------------------------
static std::mutex g_mu;
static std::map<std::tuple<int, int, int, int, int>, cufftHandle> g_plans;

// kind: 0 R2C, 1 C2R (contiguous rows of length n, batch rows); 2 D2Z strided columns, 3 Z2D strided columns
static int get_plan(int kind, int n, int batch, int stride, int dev, cufftHandle* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto key = std::make_tuple(kind, n, batch, stride, dev);
  auto it = g_plans.find(key);
  if (it != g_plans.end()) {
    *out = it->second;
    return MAUA_OK;
  }
  cufftHandle h;
  cufftResult r;
  int nn[1] = {n};
  if (kind == 0) {
    r = cufftPlan1d(&h, n, CUFFT_R2C, batch);
  } else if (kind == 1) {
    r = cufftPlan1d(&h, n, CUFFT_C2R, batch);
  } else if (kind == 2) {  // input real [n][stride] column-wise, output complex [n/2+1][stride]
    int inembed[1] = {n}, onembed[1] = {n / 2 + 1};
    r = cufftPlanMany(&h, 1, nn, inembed, stride, 1, onembed, stride, 1, CUFFT_D2Z, batch);
  } else {
    int inembed[1] = {n / 2 + 1}, onembed[1] = {n};
    r = cufftPlanMany(&h, 1, nn, inembed, stride, 1, onembed, stride, 1, CUFFT_Z2D, batch);
  }
  if (r != CUFFT_SUCCESS) {
    set_error("cuFFT plan creation failed (kind %d n %d batch %d): %d", kind, n, batch, (int)r);
    return MAUA_E_CUDA;
  }
  g_plans[key] = h;
  *out = h;
  return MAUA_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// STFT framing / ISTFT overlap-add
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float hann_periodic(int i, int n) { return 0.5f - 0.5f * cospif(2.f * i / n); }

__global__ void frame_kernel(const float* __restrict__ y, long long n, float* __restrict__ frames, int n_fft, int hop,
                             long long total) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int k = (int)(i % n_fft);
    const long long t = i / n_fft;
    long long p = t * hop + k - n_fft / 2;  // centred; reflect padding (numpy 'reflect': no edge repeat)
    if (p < 0) p = -p;
    if (p >= n) p = 2 * (n - 1) - p;
    if (p < 0) p = 0;
    frames[i] = y[p] * hann_periodic(k, n_fft);
  }
}

__global__ void overlap_add_kernel(const float* __restrict__ frames, float* __restrict__ y, long long n, int n_fft,
                                   int hop, int T) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const long long q = i + n_fft / 2;  // position in the padded signal
    long long t_hi = q / hop;
    if (t_hi > T - 1) t_hi = T - 1;
    float acc = 0.f, wss = 0.f;
    for (long long t = t_hi; t >= 0; --t) {
      const long long k = q - t * hop;
      if (k >= n_fft) break;
      const float w = hann_periodic((int)k, n_fft);
      acc = fmaf(frames[t * n_fft + k], w, acc);
      wss = fmaf(w, w, wss);
    }
    acc *= 1.f / n_fft;  // cuFFT C2R is unnormalised
    y[i] = (wss > 1.17549435e-38f) ? acc / wss : acc;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// HPSS: median filters (31 along time / frequency, scipy 'reflect' boundary) + soft mask
// ---------------------------------------------------------------------------------------------------------------
__global__ void magnitude_kernel(const float2* __restrict__ S, float* __restrict__ mag, long long n) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float2 v = S[i];
    mag[i] = sqrtf(v.x * v.x + v.y * v.y);
  }
}

__device__ __forceinline__ int reflect_sym(int i, int n) {  // d c b a | a b c d | d c b a
  if (n == 1) return 0;
  const int period = 2 * n;
  i %= period;
  if (i < 0) i += period;
  return i < n ? i : period - 1 - i;
}

template <int K>
__device__ __forceinline__ float median_of(const float* v) {
  // rank by counting (ties broken by index): O(K^2) but branch-free and register-resident
  float med = v[0];
#pragma unroll
  for (int i = 0; i < K; ++i) {
    int rank = 0;
    const float vi = v[i];
#pragma unroll
    for (int j = 0; j < K; ++j) rank += (v[j] < vi) || (v[j] == vi && j < i);
    if (rank == K / 2) med = vi;
  }
  return med;
}

template <int K>
__global__ void __launch_bounds__(128) hpss_kernel(const float2* __restrict__ S, const float* __restrict__ mag,
                                                   float2* __restrict__ out, int T, int F, float margin, float power,
                                                   int which) {
  const long long total = (long long)T * F;
  for (long long i = blockIdx.x * 128LL + threadIdx.x; i < total; i += (long long)gridDim.x * 128) {
    const int f = (int)(i % F);
    const int t = (int)(i / F);
    float v[K];
#pragma unroll
    for (int j = 0; j < K; ++j) v[j] = mag[(long long)reflect_sym(t + j - K / 2, T) * F + f];
    const float harm = median_of<K>(v);
#pragma unroll
    for (int j = 0; j < K; ++j) v[j] = mag[(long long)t * F + reflect_sym(f + j - K / 2, F)];
    const float perc = median_of<K>(v);
    const float X = which == 0 ? harm : perc;
    const float R = (which == 0 ? perc : harm) * margin;
    const float Z = fmaxf(X, R);
    float mask = 0.f;
    if (Z >= 1.17549435e-38f) {
      const float a = powf(X / Z, power), b = powf(R / Z, power);
      mask = a / (a + b);
    }
    const float2 s = S[i];
    out[i] = make_float2(s.x * mask, s.y * mask);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// filterbank projection: out[t][m] = sum_f w(S[t][f]) * fb[m][f],  w = |.|^2 (power) ; one block per frame
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) filterbank_kernel(const float2* __restrict__ S, const float* __restrict__ fb,
                                                         float* __restrict__ out, int F, int M) {
  extern __shared__ float pw[];
  const int t = blockIdx.x;
  for (int f = threadIdx.x; f < F; f += 256) {
    const float2 v = S[(long long)t * F + f];
    pw[f] = v.x * v.x + v.y * v.y;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int m = warp; m < M; m += 8) {
    float acc = 0.f;
    for (int f = lane; f < F; f += 32) acc = fmaf(pw[f], __ldg(fb + (long long)m * F + f), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[(long long)t * M + m] = acc;
  }
}

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  int* a = reinterpret_cast<int*>(addr);
  int old = *a;
  while (__int_as_float(old) < v) {
    const int assumed = old;
    old = atomicCAS(a, assumed, __float_as_int(v));
    if (old == assumed) break;
  }
}

__global__ void to_db_max_kernel(float* __restrict__ x, long long n, float amin, float* __restrict__ gmax) {
  float m = -3.0e38f;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float d = 10.f * log10f(fmaxf(amin, x[i]));
    x[i] = d;
    m = fmaxf(m, d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomic_max_float(gmax, m);
}

// env[t] = mean_m max(0, db[t][m] - db[t-1][m]) shifted right by `pad`, db clamped at gmax - top_db
__global__ void onset_env_kernel(const float* __restrict__ db, const float* __restrict__ gmax, float top_db,
                                 float* __restrict__ env, int T, int M, int pad) {
  const int t = blockIdx.x * blockDim.y + threadIdx.y;  // output index
  if (t >= T) return;
  const int src = t - pad + 1;                           // env_unpadded[src-1] = diff(db[src], db[src-1])
  float acc = 0.f;
  if (src >= 1 && src < T) {
    const float floor_db = *gmax - top_db;
    for (int m = threadIdx.x; m < M; m += 32) {
      const float a = fmaxf(db[(long long)src * M + m], floor_db), b = fmaxf(db[(long long)(src - 1) * M + m], floor_db);
      acc += fmaxf(0.f, a - b);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (threadIdx.x == 0) env[t] = acc / M;
}

__global__ void rms_kernel(const float2* __restrict__ S, float* __restrict__ r, int T, int F, int n_fft) {
  const int t = blockIdx.x;
  float acc = 0.f;
  for (int f = threadIdx.x; f < F; f += 256) {
    const float2 v = S[(long long)t * F + f];
    const float p = v.x * v.x + v.y * v.y;
    acc += (f == 0 || f == F - 1) ? 0.5f * p : p;
  }
  __shared__ float red[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    r[t] = sqrtf(2.f * s / ((float)n_fft * (float)n_fft));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// CENS post-processing and cosine k-NN median filter on chroma [T][12]
// ---------------------------------------------------------------------------------------------------------------
__global__ void cens_quant_kernel(const float* __restrict__ raw, float* __restrict__ q, int T, int C) {
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= T) return;
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += fabsf(raw[(long long)t * C + c]);
  s = fmaxf(s, 1.17549435e-38f);
  for (int c = 0; c < C; ++c) {
    const float v = raw[(long long)t * C + c] / s;
    q[(long long)t * C + c] = 0.25f * ((v > 0.4f) + (v > 0.2f) + (v > 0.1f) + (v > 0.05f));
  }
}

// (C <= 16 accumulators live in registers: the loops over c are fully unrolled and guarded, a runtime-indexed acc[] went to
//  local memory: 36 STL in the round-1 SASS)
__global__ void cens_smooth_kernel(const float* __restrict__ q, float* __restrict__ out, int T, int C, int win_len) {
  // hann(win_len, symmetric) / sum, 'same' convolution with zero boundary, then L2 normalisation over C
  const int t = blockIdx.x * 256 + threadIdx.x;
  if (t >= T) return;
  float acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = 0.f;
  float wsum = 0.f;
  for (int j = 0; j < win_len; ++j) wsum += 0.5f - 0.5f * cospif(2.f * j / (win_len - 1));
  for (int j = 0; j < win_len; ++j) {
    const int tt = t + j - win_len / 2;
    if (tt < 0 || tt >= T) continue;
    const float w = (0.5f - 0.5f * cospif(2.f * j / (win_len - 1))) / wsum;
#pragma unroll
    for (int c = 0; c < 16; ++c)
      if (c < C) acc[c] = fmaf(q[(long long)tt * C + c], w, acc[c]);
  }
  float n2 = 0.f;
#pragma unroll
  for (int c = 0; c < 16; ++c)
    if (c < C) n2 = fmaf(acc[c], acc[c], n2);
  n2 = fmaxf(sqrtf(n2), 1.17549435e-38f);
#pragma unroll
  for (int c = 0; c < 16; ++c)
    if (c < C) out[(long long)t * C + c] = acc[c] / n2;
}

// One block per frame i: cosine distances to every other frame into scratch, k-th smallest by bisection on the
// (monotone) float bit pattern, then per-bin median over the selected neighbours (rank counting), out = min(x, med).
__global__ void __launch_bounds__(256) nn_filter_kernel(const float* __restrict__ X, float* __restrict__ out,
                                                        float* __restrict__ scratch, int T, int C, int k) {
  __shared__ float xi[16];
  __shared__ float inorm;
  __shared__ int cnt;
  __shared__ float nb[16][512];  // neighbours' features (k <= 512)
  __shared__ int nsel;
  float* dist = scratch + (long long)blockIdx.x * T;
  for (int i = blockIdx.x; i < T; i += gridDim.x) {
    __syncthreads();
    if (threadIdx.x < C) xi[threadIdx.x] = X[(long long)i * C + threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
      for (int c = 0; c < C; ++c) s = fmaf(xi[c], xi[c], s);
      inorm = rsqrtf(fmaxf(s, 1e-60f));
    }
    __syncthreads();
    for (int j = threadIdx.x; j < T; j += 256) {
      float dot = 0.f, s = 0.f;
      for (int c = 0; c < C; ++c) {
        const float v = X[(long long)j * C + c];
        dot = fmaf(v, xi[c], dot);
        s = fmaf(v, v, s);
      }
      float d = 1.f - dot * inorm * rsqrtf(fmaxf(s, 1e-60f));
      d = fmaxf(d, 0.f);
      dist[j] = (j == i) ? 3.0e38f : d;
    }
    __syncthreads();
    // bisection over the non-negative float bit patterns: smallest thr with count(dist <= thr) >= k
    unsigned lo = 0u, hi = 0x7f7fffffu;
    while (lo < hi) {
      const unsigned mid = lo + (hi - lo) / 2;
      if (threadIdx.x == 0) cnt = 0;
      __syncthreads();
      int c_local = 0;
      for (int j = threadIdx.x; j < T; j += 256) c_local += (__float_as_uint(dist[j]) <= mid);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) c_local += __shfl_xor_sync(0xffffffffu, c_local, o);
      if ((threadIdx.x & 31) == 0) atomicAdd(&cnt, c_local);
      __syncthreads();
      const int total = cnt;
      __syncthreads();
      if (total >= k) hi = mid; else lo = mid + 1;
    }
    if (threadIdx.x == 0) nsel = 0;
    __syncthreads();
    for (int j = threadIdx.x; j < T; j += 256) {
      if (__float_as_uint(dist[j]) <= lo) {
        const int slot = atomicAdd(&nsel, 1);
        if (slot < 512)
          for (int c = 0; c < C; ++c) nb[c][slot] = X[(long long)j * C + c];
      }
    }
    __syncthreads();
    const int n = min(nsel, 512);
    // median per bin: average of the two middle order statistics when n is even (numpy.median)
    for (int c = threadIdx.x >> 5; c < C; c += 8) {
      const int lane = threadIdx.x & 31;
      float lo_v = 0.f, hi_v = 0.f;
      const int r_lo = (n - 1) / 2, r_hi = n / 2;
      for (int a = lane; a < n; a += 32) {
        const float va = nb[c][a];
        int rank = 0;
        for (int b = 0; b < n; ++b) rank += (nb[c][b] < va) || (nb[c][b] == va && b < a);
        if (rank == r_lo) lo_v = va;
        if (rank == r_hi) hi_v = va;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo_v += __shfl_xor_sync(0xffffffffu, lo_v, o);  // exactly one lane holds a non-zero candidate (or zeros)
        hi_v += __shfl_xor_sync(0xffffffffu, hi_v, o);
      }
      if (lane == 0) out[(long long)i * C + c] = fminf(xi[c], 0.5f * (lo_v + hi_v));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Fourier resampling along time (scipy.signal.resample), columns = channels, double precision
// ---------------------------------------------------------------------------------------------------------------
__global__ void f2d_kernel(const float* __restrict__ x, double* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) y[i] = (double)x[i];
}
__global__ void resample_spectrum_kernel(const double2* __restrict__ X, double2* __restrict__ Y, int n_in, int n_out,
                                         int C) {
  const int fo = n_out / 2 + 1, fi = n_in / 2 + 1;
  const int N = n_in < n_out ? n_in : n_out;
  const double scale = 1.0 / (double)n_in;  // Z2D is unnormalised: y = irfft(Y) * (n_out / n_in) with 1/n_out inside
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < (long long)fo * C; i += (long long)gridDim.x * 256) {
    const int k = (int)(i / C), c = (int)(i % C);
    double2 v = make_double2(0.0, 0.0);
    if (k < N / 2 + 1 && k < fi) {
      v = X[(long long)k * C + c];
      if ((N % 2 == 0) && k == N / 2) {
        if (n_out < n_in) { v.x *= 2.0; v.y *= 2.0; }
        else if (n_in < n_out) { v.x *= 0.5; v.y *= 0.5; }
      }
    }
    Y[i] = make_double2(v.x * scale, v.y * scale);
  }
}
__global__ void d2f_clip_kernel(const double* __restrict__ x, float* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += (long long)gridDim.x * 256) y[i] = (float)x[i];
}

// ---------------------------------------------------------------------------------------------------------------
// time-axis gaussian filter (signal.py:319-368), x [T][N] -> y [T][N], circular
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gaussian_filter_kernel(const float* __restrict__ x, float* __restrict__ y, int T,
                                                              long long N, float sigma, int radius, int causal_mode,
                                                              float causal) {
  extern __shared__ float g[];  // 2*radius+1 taps
  __shared__ float gsum;
  const int nt = 2 * radius + 1;
  for (int j = threadIdx.x; j < nt; j += 256) {
    const float k = (float)(j - radius);
    float v = expf(-0.5f / (sigma * sigma) * k * k);
    if (j > radius && causal_mode == 1) v *= causal;
    if (j > radius && causal_mode == 2) v = 0.f;
    g[j] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int j = 0; j < nt; ++j) s += g[j];
    gsum = s;
  }
  __syncthreads();
  const float inv = 1.f / gsum;
  const long long total = (long long)T * N;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long n = i % N;
    const int t = (int)(i / N);
    float acc = 0.f;
    if (radius <= T) {
      int tt = t - radius;
      tt %= T;
      if (tt < 0) tt += T;
      for (int j = 0; j < nt; ++j) {
        acc = fmaf(g[j] * inv, __ldg(x + (long long)tt * N + n), acc);
        if (++tt == T) tt = 0;
      }
    } else {  // [zeros(r-T), x, x, x, zeros(r-T)]
      for (int j = 0; j < nt; ++j) {
        const int q = t + j - (radius - T);
        if (q < 0 || q >= 3 * T) continue;
        acc = fmaf(g[j] * inv, __ldg(x + (long long)(q % T) * N + n), acc);
      }
    }
    y[i] = acc;
  }
}

// Register-tiled version for radius <= T (the common case): one thread = one series n and TT consecutive output frames.
// Every input sample is loaded ONCE and feeds TT accumulators; the 2*TT-1 taps a group of TT inputs needs sit in
// registers (TT new taps from shared memory per TT*TT FMAs).  The one-output-per-thread kernel above re-reads every input
// 2r+1 times: ncu on the sigma = 128 noise filter ([900, 256*256]) showed 74 GB of DRAM reads for a 236 MB tensor, 35 ms.
template <int TT>
__global__ void __launch_bounds__(256) gaussian_filter_tiled_kernel(const float* __restrict__ x, float* __restrict__ y, int T,
                                                                    long long N, float sigma, int radius, int causal_mode,
                                                                    float causal) {
  extern __shared__ float g[];  // TT-1 zeros | 2*radius+1 normalised taps | 2*TT zeros
  __shared__ float gsum;
  const int nt = 2 * radius + 1;
  float* taps = g + (TT - 1);
  for (int j = threadIdx.x; j < nt + 3 * TT; j += 256) g[j] = 0.f;
  __syncthreads();
  for (int j = threadIdx.x; j < nt; j += 256) {
    const float k = (float)(j - radius);
    float v = expf(-0.5f / (sigma * sigma) * k * k);
    if (j > radius && causal_mode == 1) v *= causal;
    if (j > radius && causal_mode == 2) v = 0.f;
    taps[j] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int j = 0; j < nt; ++j) s += taps[j];   // same summation order as the reference kernel above
    gsum = s;
  }
  __syncthreads();
  const float inv = 1.f / gsum;
  for (int j = threadIdx.x; j < nt; j += 256) taps[j] *= inv;
  __syncthreads();

  const long long n = blockIdx.x * 256LL + threadIdx.x;
  const int t0 = blockIdx.y * TT;
  if (n >= N) return;
  float acc[TT];
#pragma unroll
  for (int k = 0; k < TT; ++k) acc[k] = 0.f;
  // output t0+k = sum_j taps[j] * x[t0 + k - radius + j]; walk the inputs p = t0 - radius + q, q = 0 .. nt + TT - 2:
  // input q feeds output k with tap index j = q - k.  Inputs are taken in groups of TT (q = q0 + jj); a group needs the
  // taps q0 - (TT-1) .. q0 + TT - 1, read from the zero-padded table (indices below 0 / above nt-1 are zero).
  int tt = (t0 - radius) % T;
  if (tt < 0) tt += T;
  const int n_in = nt + TT - 1;
  for (int q0 = 0; q0 < n_in; q0 += TT) {
    float w[2 * TT - 1];
#pragma unroll
    for (int i = 0; i < 2 * TT - 1; ++i) w[i] = g[q0 + i];   // g[q0 + i] = taps[q0 + i - (TT-1)]
#pragma unroll
    for (int jj = 0; jj < TT; ++jj) {
      const float v = (q0 + jj < n_in) ? __ldg(x + (long long)tt * N + n) : 0.f;
      if (++tt == T) tt = 0;
#pragma unroll
      for (int k = 0; k < TT; ++k) acc[k] = fmaf(w[jj - k + TT - 1], v, acc[k]);   // tap (q0 + jj) - k
    }
  }
#pragma unroll
  for (int k = 0; k < TT; ++k)
    if (t0 + k < T) y[(long long)(t0 + k) * N + n] = acc[k];
}

// ---------------------------------------------------------------------------------------------------------------
// percentile_clip (signal.py:273-292): single block; peaks -> bitonic sort in shared memory -> k-th -> clamp -> /max
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) percentile_clip_kernel(const float* __restrict__ x, float* __restrict__ y, int T,
                                                               float p, int cap) {
  extern __shared__ float pk[];  // cap (power of two) floats
  __shared__ int npk;
  __shared__ float thr_s, max_s;
  if (threadIdx.x == 0) { npk = 0; max_s = 0.f; }
  for (int i = threadIdx.x; i < cap; i += 1024) pk[i] = 3.0e38f;
  __syncthreads();
  for (int i = threadIdx.x; i < T; i += 1024) {
    const float v = x[i];
    const float a = x[i + 1 < T ? i + 1 : T - 1], b = x[i > 0 ? i - 1 : 0];
    if (v > a && v > b) pk[atomicAdd(&npk, 1)] = v;
  }
  __syncthreads();
  for (int size = 2; size <= cap; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < cap; i += 1024) {
        const int j = i ^ stride;
        if (j > i) {
          const bool up = (i & size) == 0;
          const float a = pk[i], b = pk[j];
          if ((a > b) == up) { pk[i] = b; pk[j] = a; }
        }
      }
      __syncthreads();
    }
  if (threadIdx.x == 0) {
    const int n = npk;
    int k = 1 + (int)rint(0.01 * (double)p * (double)(n - 1));  // python round() == rint (half to even)
    if (k < 1) k = 1;
    if (k > n) k = n;
    thr_s = n > 0 ? pk[k - 1] : 3.0e38f;
  }
  __syncthreads();
  const float thr = thr_s;
  float m = 0.f;
  for (int i = threadIdx.x; i < T; i += 1024) m = fmaxf(m, fminf(fmaxf(x[i], 0.f), thr));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomic_max_float(&max_s, m);
  __syncthreads();
  const float mx = max_s;
  for (int i = threadIdx.x; i < T; i += 1024) y[i] = fminf(fmaxf(x[i], 0.f), thr) / mx;
}

__global__ void pow_kernel(float* __restrict__ x, int n, float p) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i < n) x[i] = powf(x[i], p);
}

__global__ void minmax_clip_kernel(const float* __restrict__ ref, int n_ref, float* __restrict__ x, int n) {
  // np.clip(resampled, ref.min(), ref.max()) — single block of 256 threads
  __shared__ float wmin[8], wmax[8];
  float lo = 3.0e38f, hi = -3.0e38f;
  for (int i = threadIdx.x; i < n_ref; i += 256) { lo = fminf(lo, ref[i]); hi = fmaxf(hi, ref[i]); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { wmin[threadIdx.x >> 5] = lo; wmax[threadIdx.x >> 5] = hi; }
  __syncthreads();
  lo = wmin[0]; hi = wmax[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) { lo = fminf(lo, wmin[i]); hi = fmaxf(hi, wmax[i]); }
  for (int i = threadIdx.x; i < n; i += 256) x[i] = fminf(fmaxf(x[i], lo), hi);
}

// sequential IIR (scipy.signal.sosfilt, direct form II transposed), double precision, one thread
template <int NS>
__global__ void sosfilt_kernel(const float* __restrict__ x, float* __restrict__ y, long long n,
                               const double* __restrict__ sos) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  // section state and coefficients in registers (NS is a compile-time constant: no local-memory arrays)
  double z0[NS], z1[NS], c0[NS], c1[NS], c2[NS], c4[NS], c5[NS];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    z0[s] = z1[s] = 0.0;
    const double* c = sos + 6 * s;  // b0 b1 b2 a0 a1 a2 (a0 == 1)
    c0[s] = c[0]; c1[s] = c[1]; c2[s] = c[2]; c4[s] = c[4]; c5[s] = c[5];
  }
  for (long long i = 0; i < n; ++i) {
    double v = (double)x[i];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      const double o = c0[s] * v + z0[s];
      z0[s] = c1[s] * v - c4[s] * o + z1[s];
      z1[s] = c2[s] * v - c5[s] * o;
      v = o;
    }
    y[i] = (float)v;
  }
}

// latents: base[t][e] = sum_n chroma[t][n] * sel[n][e]      (latent.py:15-26)
__global__ void chroma_weight_kernel(const float* __restrict__ chroma, const float* __restrict__ sel,
                                     float* __restrict__ out, int T, int NS, long long E) {
  const long long total = (long long)T * E;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const long long e = i % E;
    const int t = (int)(i / E);
    float acc = 0.f;
    for (int n = 0; n < NS; ++n) acc = fmaf(__ldg(chroma + (long long)t * NS + n), __ldg(sel + (long long)n * E + e), acc);
    out[i] = acc;
  }
}
// x[t][e] = w[t] * a[e] + (1 - w[t]) * x[t][e]               (examples/default.py:20-21)
__global__ void envelope_blend_kernel(float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ a,
                                      int T, long long E) {
  const long long total = (long long)T * E;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const float wt = __ldg(w + i / E);
    x[i] = wt * __ldg(a + i % E) + (1.f - wt) * x[i];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// madmom-flavoured onsets (the reference's DEFAULT `type="mm"`, audioreactive/signal.py:52-67; madmom is un-vendored:
// restated from its published algorithm — FramedSignal / ShortTimeFourierTransform(circular_shift) /
// LogarithmicFilterbank / features.onsets.{spectral_diff, spectral_flux, superflux, complex_flux,
// modified_kullback_leibler}; parity unpinned)
// ---------------------------------------------------------------------------------------------------------------
// FramedSignal(frame_size, hop, origin 0): frame t = samples [t*hop - n_fft/2, +n_fft), zeros outside; symmetric
// np.hanning window; circular shift by n_fft/2 so that phases are referenced to the frame centre.
__global__ void frame_mm_kernel(const float* __restrict__ y, long long n, float* __restrict__ frames, int n_fft, int hop,
                                long long total) {
  const int half = n_fft / 2;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int k = (int)(i % n_fft);             // position in the SHIFTED frame
    const long long t = i / n_fft;
    const int ks = k < half ? k + half : k - half;  // position in the windowed frame
    const long long p = t * hop + ks - half;
    const float w = 0.5f - 0.5f * cospif(2.f * ks / (n_fft - 1));
    frames[i] = (p >= 0 && p < n) ? y[p] * w : 0.f;
  }
}

// per frame: filtered magnitude spectrogram row and |local group delay| / pi row
__global__ void __launch_bounds__(256) mm_filt_lgd_kernel(const float2* __restrict__ S, const float* __restrict__ fb,
                                                          float* __restrict__ filt, float* __restrict__ lgd, int Fs,
                                                          int F, int NB) {
  extern __shared__ float sm[];  // mag[F], phase[F]
  float* mag = sm;
  float* ph = sm + F;
  const int t = blockIdx.x;
  for (int f = threadIdx.x; f < F; f += 256) {
    const float2 v = S[(long long)t * Fs + f];
    mag[f] = sqrtf(v.x * v.x + v.y * v.y);
    ph[f] = atan2f(v.y, v.x);
  }
  __syncthreads();
  for (int f = threadIdx.x; f < F; f += 256) {
    float g = 0.f;
    if (f + 1 < F) {
      // np.unwrap along frequency, then lgd[f] = up[f] - up[f+1]: minus the wrapped phase step
      const float dd = ph[f + 1] - ph[f];
      float dm = dd + kPi;
      dm = dm - 2.f * kPi * floorf(dm / (2.f * kPi)) - kPi;
      if (dm == -kPi && dd > 0.f) dm = kPi;
      g = fabsf(fabsf(dd) < kPi ? dd : dm) / kPi;
    }
    lgd[(long long)t * F + f] = g;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = warp; b < NB; b += 8) {
    float acc = 0.f;
    for (int f = lane; f < F; f += 32) acc = fmaf(mag[f], __ldg(fb + (long long)b * F + f), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) filt[(long long)t * NB + b] = acc;
  }
}

// per frame: sum of the five onset detection functions
__global__ void __launch_bounds__(128) mm_onset_kernel(const float* __restrict__ filt, const float* __restrict__ lgd,
                                                       const int* __restrict__ band_lo, const int* __restrict__ band_hi,
                                                       float* __restrict__ onset, int T, int F, int NB, int diff_frames,
                                                       float eps) {
  const int t = blockIdx.x;
  float sd = 0.f, sf = 0.f, su = 0.f, cf = 0.f, kl = 0.f;
  const int tp = t - diff_frames;
  for (int b = threadIdx.x; b < NB; b += 128) {
    const float cur = filt[(long long)t * NB + b];
    if (tp >= 0) {
      const float* prev = filt + (long long)tp * NB;
      const float d = fmaxf(cur - prev[b], 0.f);
      sd = fmaf(d, d, sd);
      sf += d;
      // SuperFlux: previous frame max-filtered over 3 bands (scipy maximum_filter, 'reflect' edges)
      const float pm = fmaxf(prev[b], fmaxf(prev[b > 0 ? b - 1 : 0], prev[b + 1 < NB ? b + 1 : NB - 1]));
      const float dm = fmaxf(cur - pm, 0.f);
      su += dm;
      // ComplexFlux mask: min over the band's bins (+1 neighbour each side) of the 3-frame temporal max of |lgd|
      float mask = 3.0e38f;
      const int tm = t > 0 ? t - 1 : 0, tn = t + 1 < T ? t + 1 : T - 1;
      for (int f = band_lo[b]; f < band_hi[b]; ++f) {
        const float g = fmaxf(lgd[(long long)t * F + f], fmaxf(lgd[(long long)tm * F + f], lgd[(long long)tn * F + f]));
        mask = fminf(mask, g);
      }
      cf = fmaf(dm, mask, cf);
    }
    if (t >= 1) kl += log1pf(cur / (filt[(long long)(t - 1) * NB + b] + eps));
  }
  __shared__ float red[5][4];
  float v[5] = {sd, sf, su, cf, kl};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    if (lane == 0) red[i][warp] = v[i];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) tot[i] = red[i][0] + red[i][1] + red[i][2] + red[i][3];
    onset[t] = tot[0] + tot[1] + tot[2] + tot[3] + tot[4] / (float)NB;
  }
}

static inline unsigned nblocks(long long work, int per = 256, long long cap = 148LL * 16) {
  long long b = (work + per - 1) / per;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace audio
}  // namespace maua

using namespace maua;
using namespace maua::audio;

#define MAUA_CHECK_FFT(expr)                                         \
  do {                                                               \
    cufftResult _r = (expr);                                         \
    if (_r != CUFFT_SUCCESS) {                                       \
      set_error("%s failed: cufftResult %d", #expr, (int)_r);        \
      return MAUA_E_CUDA;                                            \
    }                                                                \
  } while (0)

static int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return d;
}

extern "C" int maua_audio_stft_f32(const float* y, long long n, float* spec, float* frames_ws, int n_fft, int hop,
                                   int n_frames, void* stream) {
  MAUA_CHECK_ARG(y && spec && frames_ws && n >= 2 && n_fft >= 2 && hop >= 1 && n_frames >= 1, "stft: bad arguments");
  cudaStream_t st = as_stream(stream);
  const long long total = (long long)n_frames * n_fft;
  frame_kernel<<<nblocks(total), 256, 0, st>>>(y, n, frames_ws, n_fft, hop, total);
  MAUA_CHECK_LAUNCH("stft(frame)");
  cufftHandle plan;
  int rc = get_plan(0, n_fft, n_frames, 1, current_device(), &plan);
  if (rc) return rc;
  MAUA_CHECK_FFT(cufftSetStream(plan, st));
  MAUA_CHECK_FFT(cufftExecR2C(plan, frames_ws, reinterpret_cast<cufftComplex*>(spec)));
  count_launch();
  return MAUA_OK;
}

extern "C" int maua_audio_istft_f32(float* spec, float* y, long long n, float* frames_ws, int n_fft, int hop,
                                    int n_frames, void* stream) {
  MAUA_CHECK_ARG(y && spec && frames_ws && n >= 1 && n_fft >= 2 && hop >= 1 && n_frames >= 1, "istft: bad arguments");
  cudaStream_t st = as_stream(stream);
  cufftHandle plan;
  int rc = get_plan(1, n_fft, n_frames, 1, current_device(), &plan);
  if (rc) return rc;
  MAUA_CHECK_FFT(cufftSetStream(plan, st));
  MAUA_CHECK_FFT(cufftExecC2R(plan, reinterpret_cast<cufftComplex*>(spec), frames_ws));  // may clobber `spec`
  count_launch();
  overlap_add_kernel<<<nblocks(n), 256, 0, st>>>(frames_ws, y, n, n_fft, hop, n_frames);
  MAUA_CHECK_LAUNCH("istft(overlap_add)");
  return MAUA_OK;
}

extern "C" int maua_audio_hpss_f32(const float* spec, float* spec_out, float* mag_ws, int n_frames, int n_bins,
                                   float margin, float power, int which, void* stream) {
  MAUA_CHECK_ARG(spec && spec_out && mag_ws && n_frames >= 1 && n_bins >= 1 && (which == 0 || which == 1),
                 "hpss: bad arguments");
  cudaStream_t st = as_stream(stream);
  const long long n = (long long)n_frames * n_bins;
  magnitude_kernel<<<nblocks(n), 256, 0, st>>>(reinterpret_cast<const float2*>(spec), mag_ws, n);
  MAUA_CHECK_LAUNCH("hpss(magnitude)");
  hpss_kernel<31><<<nblocks(n, 128, 148LL * 64), 128, 0, st>>>(reinterpret_cast<const float2*>(spec), mag_ws,
                                                              reinterpret_cast<float2*>(spec_out), n_frames, n_bins,
                                                              margin, power, which);
  MAUA_CHECK_LAUNCH("hpss(median+mask)");
  return MAUA_OK;
}

extern "C" int maua_audio_filterbank_f32(const float* spec, const float* fb, float* out, int n_frames, int n_bins,
                                         int n_filters, void* stream) {
  MAUA_CHECK_ARG(spec && fb && out && n_frames >= 1 && n_bins >= 1 && n_filters >= 1 && n_bins <= 8192,
                 "filterbank: bad arguments");
  filterbank_kernel<<<n_frames, 256, n_bins * sizeof(float), as_stream(stream)>>>(
      reinterpret_cast<const float2*>(spec), fb, out, n_bins, n_filters);
  MAUA_CHECK_LAUNCH("filterbank");
  return MAUA_OK;
}

extern "C" int maua_audio_onset_env_f32(float* mel, float* env, float* scalar_ws, int n_frames, int n_mels, int pad,
                                        float amin, float top_db, void* stream) {
  MAUA_CHECK_ARG(mel && env && scalar_ws && n_frames >= 2 && n_mels >= 1 && pad >= 0, "onset_env: bad arguments");
  cudaStream_t st = as_stream(stream);
  const float neg = -3.0e38f;
  MAUA_CHECK_CUDA(cudaMemcpyAsync(scalar_ws, &neg, sizeof(float), cudaMemcpyHostToDevice, st));
  const long long n = (long long)n_frames * n_mels;
  to_db_max_kernel<<<nblocks(n), 256, 0, st>>>(mel, n, amin, scalar_ws);
  MAUA_CHECK_LAUNCH("onset_env(db)");
  onset_env_kernel<<<ceil_div(n_frames, 8), dim3(32, 8), 0, st>>>(mel, scalar_ws, top_db, env, n_frames, n_mels, pad);
  MAUA_CHECK_LAUNCH("onset_env(diff)");
  return MAUA_OK;
}

extern "C" int maua_audio_stft_mm_f32(const float* y, long long n, float* spec, float* frames_ws, int n_fft, int hop,
                                      int n_frames, void* stream) {
  MAUA_CHECK_ARG(y && spec && frames_ws && n >= 1 && n_fft >= 4 && n_fft % 2 == 0 && hop >= 1 && n_frames >= 1,
                 "stft_mm: bad arguments");
  cudaStream_t st = as_stream(stream);
  const long long total = (long long)n_frames * n_fft;
  frame_mm_kernel<<<nblocks(total), 256, 0, st>>>(y, n, frames_ws, n_fft, hop, total);
  MAUA_CHECK_LAUNCH("stft_mm(frame)");
  cufftHandle plan;
  int rc = get_plan(0, n_fft, n_frames, 1, current_device(), &plan);
  if (rc) return rc;
  MAUA_CHECK_FFT(cufftSetStream(plan, st));
  MAUA_CHECK_FFT(cufftExecR2C(plan, frames_ws, reinterpret_cast<cufftComplex*>(spec)));
  count_launch();
  return MAUA_OK;
}

extern "C" int maua_audio_onsets_mm_f32(const float* spec, const float* fb, const int* band_lo, const int* band_hi,
                                        float* onset, float* filt_ws, float* lgd_ws, int n_frames, int spec_stride,
                                        int n_bins, int n_bands, int diff_frames, void* stream) {
  MAUA_CHECK_ARG(spec && fb && band_lo && band_hi && onset && filt_ws && lgd_ws, "onsets_mm: null pointers");
  MAUA_CHECK_ARG(n_frames >= 1 && n_bins >= 2 && n_bins <= spec_stride && n_bins <= 8192 && n_bands >= 1 &&
                     diff_frames >= 1,
                 "onsets_mm: bad shape");
  cudaStream_t st = as_stream(stream);
  mm_filt_lgd_kernel<<<n_frames, 256, 2 * (size_t)n_bins * sizeof(float), st>>>(
      reinterpret_cast<const float2*>(spec), fb, filt_ws, lgd_ws, spec_stride, n_bins, n_bands);
  MAUA_CHECK_LAUNCH("onsets_mm(filter+lgd)");
  mm_onset_kernel<<<n_frames, 128, 0, st>>>(filt_ws, lgd_ws, band_lo, band_hi, onset, n_frames, n_bins, n_bands,
                                            diff_frames, 2.220446049250313e-16f);
  MAUA_CHECK_LAUNCH("onsets_mm(sum)");
  return MAUA_OK;
}

extern "C" int maua_audio_rms_f32(const float* spec, float* rms, int n_frames, int n_bins, int n_fft, void* stream) {
  MAUA_CHECK_ARG(spec && rms && n_frames >= 1 && n_bins >= 1, "rms: bad arguments");
  rms_kernel<<<n_frames, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float2*>(spec), rms, n_frames, n_bins,
                                                       n_fft);
  MAUA_CHECK_LAUNCH("rms");
  return MAUA_OK;
}

extern "C" int maua_audio_cens_f32(const float* raw, float* cens, float* ws, int n_frames, int n_chroma, int win_len,
                                   void* stream) {
  MAUA_CHECK_ARG(raw && cens && ws && n_frames >= 1 && n_chroma >= 1 && n_chroma <= 16 && win_len >= 3,
                 "cens: bad arguments");
  cudaStream_t st = as_stream(stream);
  cens_quant_kernel<<<ceil_div(n_frames, 256), 256, 0, st>>>(raw, ws, n_frames, n_chroma);
  MAUA_CHECK_LAUNCH("cens(quant)");
  cens_smooth_kernel<<<ceil_div(n_frames, 256), 256, 0, st>>>(ws, cens, n_frames, n_chroma, win_len);
  MAUA_CHECK_LAUNCH("cens(smooth)");
  return MAUA_OK;
}

extern "C" int maua_audio_nn_filter_f32(const float* x, float* out, float* scratch, int scratch_rows, int n_frames,
                                        int n_chroma, int k, void* stream) {
  MAUA_CHECK_ARG(x && out && scratch && scratch_rows >= 1 && n_frames >= 2 && n_chroma >= 1 && n_chroma <= 16,
                 "nn_filter: bad arguments");
  MAUA_CHECK_ARG(k >= 1 && k <= 512 && k < n_frames, "nn_filter: k must be in [1, 512]");
  nn_filter_kernel<<<scratch_rows, 256, 0, as_stream(stream)>>>(x, out, scratch, n_frames, n_chroma, k);
  MAUA_CHECK_LAUNCH("nn_filter");
  return MAUA_OK;
}

extern "C" int maua_resample_f32(const float* x, float* y, double* ws, int n_in, int n_out, int channels,
                                 void* stream) {
  // ws: n_in*C doubles + (n_in/2+1)*C double2 + (n_out/2+1)*C double2 + n_out*C doubles
  MAUA_CHECK_ARG(x && y && ws && n_in >= 2 && n_out >= 2 && channels >= 1, "resample: bad arguments");
  cudaStream_t st = as_stream(stream);
  const long long C = channels;
  MAUA_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "resample: workspace must be 16-byte aligned");
  double* xin = ws;
  const long long n_in_pad = ((long long)n_in * C + 1) & ~1LL;  // keep the complex buffers 16-byte aligned
  double2* X = reinterpret_cast<double2*>(xin + n_in_pad);
  double2* Y = X + (long long)(n_in / 2 + 1) * C;
  double* yout = reinterpret_cast<double*>(Y + (long long)(n_out / 2 + 1) * C);
  f2d_kernel<<<nblocks((long long)n_in * C), 256, 0, st>>>(x, xin, (long long)n_in * C);
  MAUA_CHECK_LAUNCH("resample(f2d)");
  cufftHandle fwd, inv;
  int rc;
  if ((rc = get_plan(2, n_in, channels, channels, current_device(), &fwd))) return rc;
  if ((rc = get_plan(3, n_out, channels, channels, current_device(), &inv))) return rc;
  MAUA_CHECK_FFT(cufftSetStream(fwd, st));
  MAUA_CHECK_FFT(cufftExecD2Z(fwd, xin, reinterpret_cast<cufftDoubleComplex*>(X)));
  resample_spectrum_kernel<<<nblocks((long long)(n_out / 2 + 1) * C), 256, 0, st>>>(X, Y, n_in, n_out, channels);
  MAUA_CHECK_LAUNCH("resample(spectrum)");
  MAUA_CHECK_FFT(cufftSetStream(inv, st));
  MAUA_CHECK_FFT(cufftExecZ2D(inv, reinterpret_cast<cufftDoubleComplex*>(Y), yout));
  d2f_clip_kernel<<<nblocks((long long)n_out * C), 256, 0, st>>>(yout, y, (long long)n_out * C);
  MAUA_CHECK_LAUNCH("resample(d2f)");
  count_launch(2);
  return MAUA_OK;
}

extern "C" int maua_clip_to_range_f32(const float* ref, int n_ref, float* x, int n, void* stream) {
  MAUA_CHECK_ARG(ref && x && n_ref >= 1 && n >= 1, "clip_to_range: bad arguments");
  minmax_clip_kernel<<<1, 256, 0, as_stream(stream)>>>(ref, n_ref, x, n);
  MAUA_CHECK_LAUNCH("clip_to_range");
  return MAUA_OK;
}

extern "C" int maua_gaussian_filter_f32(const float* x, float* y, int n_frames, long long inner, float sigma,
                                        float smf, int causal_mode, float causal, void* stream) {
  MAUA_CHECK_ARG(x && y && x != y && n_frames >= 1 && inner >= 1 && sigma > 0.f, "gaussian_filter: bad arguments");
  MAUA_CHECK_ARG(causal_mode >= 0 && causal_mode <= 2, "gaussian_filter: causal_mode must be 0 (none), 1 (scale), 2 (zero)");
  int radius = (int)(sigma * 4.f * smf);
  if (radius > 3 * n_frames) radius = 3 * n_frames;
  const size_t smem = (2 * (size_t)radius + 1) * sizeof(float);
  MAUA_CHECK_ARG(smem <= 200 * 1024, "gaussian_filter: radius too large");
  if (smem > 48 * 1024)
    MAUA_CHECK_CUDA(cudaFuncSetAttribute(gaussian_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  constexpr int TT = 16;
  const long long nbx = (inner + 255) / 256;
  const int nby = (n_frames + TT - 1) / TT;
  if (radius <= n_frames && radius >= 8 && nbx <= 0x7fffffffLL && nby <= 65535) {
    // register-tiled kernel: 16 output frames per thread (short filters / tiny series keep the simple kernel)
    const size_t smem_t = (2 * (size_t)radius + 1 + 3 * TT) * sizeof(float);
    if (smem_t > 48 * 1024)
      MAUA_CHECK_CUDA(cudaFuncSetAttribute(gaussian_filter_tiled_kernel<TT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem_t));
    gaussian_filter_tiled_kernel<TT><<<dim3((unsigned)nbx, (unsigned)nby), 256, smem_t, as_stream(stream)>>>(
        x, y, n_frames, inner, sigma, radius, causal_mode, causal);
    MAUA_CHECK_LAUNCH("gaussian_filter(tiled)");
    return MAUA_OK;
  }
  gaussian_filter_kernel<<<nblocks((long long)n_frames * inner, 256, 148LL * 32), 256, smem, as_stream(stream)>>>(
      x, y, n_frames, inner, sigma, radius, causal_mode, causal);
  MAUA_CHECK_LAUNCH("gaussian_filter");
  return MAUA_OK;
}

extern "C" int maua_percentile_clip_f32(const float* x, float* y, int n, float percentile, float power,
                                        void* stream) {
  MAUA_CHECK_ARG(x && y && n >= 3 && n <= 65536, "percentile_clip: n must be in [3, 65536]");
  int cap = 2;
  while (cap < (n + 1) / 2 + 1) cap <<= 1;
  const size_t smem = cap * sizeof(float);
  if (smem > 48 * 1024)
    MAUA_CHECK_CUDA(cudaFuncSetAttribute(percentile_clip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  percentile_clip_kernel<<<1, 1024, smem, as_stream(stream)>>>(x, y, n, percentile, cap);
  MAUA_CHECK_LAUNCH("percentile_clip");
  if (power != 1.f) {
    pow_kernel<<<ceil_div(n, 256), 256, 0, as_stream(stream)>>>(y, n, power);
    MAUA_CHECK_LAUNCH("percentile_clip(pow)");
  }
  return MAUA_OK;
}

extern "C" int maua_sosfilt_f32(const float* x, float* y, long long n, const double* sos, int n_sections,
                                void* stream) {
  MAUA_CHECK_ARG(x && y && sos && n >= 1 && n_sections >= 1 && n_sections <= 16, "sosfilt: bad arguments");
  cudaStream_t st = as_stream(stream);
  switch (n_sections) {   // the butter(12, bandpass) of audioreactive.rms has 12 sections; other orders are rare
#define MAUA_SOS_CASE(N) case N: sosfilt_kernel<N><<<1, 32, 0, st>>>(x, y, n, sos); break;
    MAUA_SOS_CASE(1) MAUA_SOS_CASE(2) MAUA_SOS_CASE(3) MAUA_SOS_CASE(4) MAUA_SOS_CASE(5) MAUA_SOS_CASE(6) MAUA_SOS_CASE(7)
    MAUA_SOS_CASE(8) MAUA_SOS_CASE(9) MAUA_SOS_CASE(10) MAUA_SOS_CASE(11) MAUA_SOS_CASE(12) MAUA_SOS_CASE(13)
    MAUA_SOS_CASE(14) MAUA_SOS_CASE(15) MAUA_SOS_CASE(16)
#undef MAUA_SOS_CASE
  }
  MAUA_CHECK_LAUNCH("sosfilt");
  return MAUA_OK;
}

extern "C" int maua_chroma_weight_latents_f32(const float* chroma, const float* selection, float* out, int n_frames,
                                              int n_select, long long latent_elems, void* stream) {
  MAUA_CHECK_ARG(chroma && selection && out && n_frames >= 1 && n_select >= 1 && latent_elems >= 1,
                 "chroma_weight_latents: bad arguments");
  chroma_weight_kernel<<<nblocks((long long)n_frames * latent_elems, 256, 148LL * 32), 256, 0, as_stream(stream)>>>(
      chroma, selection, out, n_frames, n_select, latent_elems);
  MAUA_CHECK_LAUNCH("chroma_weight_latents");
  return MAUA_OK;
}

extern "C" int maua_envelope_blend_f32(float* x, const float* envelope, const float* target, int n_frames,
                                       long long inner, void* stream) {
  MAUA_CHECK_ARG(x && envelope && target && n_frames >= 1 && inner >= 1, "envelope_blend: bad arguments");
  envelope_blend_kernel<<<nblocks((long long)n_frames * inner, 256, 148LL * 32), 256, 0, as_stream(stream)>>>(
      x, envelope, target, n_frames, inner);
  MAUA_CHECK_LAUNCH("envelope_blend");
  return MAUA_OK;
}
