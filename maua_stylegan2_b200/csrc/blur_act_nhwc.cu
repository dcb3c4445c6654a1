// Blur (4x4 FIR, pad (1,1)) + noise + bias + leaky-ReLU + next-layer style + bf16 split, channels-last.
//
// Second half of an upsampling StyledConv on the tensor-core path (models/stylegan2.py:238 Blur.forward,
// :262-266 NoiseInjection, op/fused_act.py:82-97): the transposed-conv kernel leaves u = d * convT(x) as fp32
// [B, 2H+1, 2W+1, C]; this kernel computes
//     o[b,oy,ox,c] = sum_{a,e<4} Kf[a][e] * u[b, oy+a-1, ox+e-1, c]        (Kf = flipped FIR, zero outside u)
//     o = o * d[b,c]                    (optional: demodulation commutes with the per-channel FIR, so the conv kernel
//                                        writes the raw phases and its epilogue needs no per-channel operand at all)
//     o = lrelu(o + nw*noise[b,oy,ox] + bias[c]) * sqrt2
// and writes what the next layer consumes: (hi, lo) bf16 NHWC of o * s_next[b,c] and/or fp32 NCHW o.
// HBM-bound: 4 B/elt in, 4 B/elt out.  One CTA = 16x16 output pixels x 32 channels; the 19x19x32 input tile
// is staged in shared memory with 128-byte-contiguous loads; each thread slides down a column of 8 outputs for
// one float4 of channels (44 LDS.128 per 8 outputs), stores are 8-byte bf16x4 per plane.
#include <cstdlib>

#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tmap.cuh"

namespace maua {

constexpr int BT = 16;          // output tile edge
constexpr int BIN = BT + 3;     // staged input edge
constexpr int BCH = 32;         // channels per CTA

__global__ void __launch_bounds__(256) blur_act_nhwc_kernel(const float* __restrict__ u, const float* __restrict__ k4,
                                                            MauaConvEpilogue ep, int ch, int hu, int wu) {
  __shared__ __align__(16) float tile[BIN * BIN * BCH];
  __shared__ float kf[16];
  const int tid = threadIdx.x;
  const int oh = hu - 1, ow = wu - 1;
  const int tiles_x = (ow + BT - 1) / BT;
  const int oy0 = (blockIdx.x / tiles_x) * BT;
  const int ox0 = (blockIdx.x % tiles_x) * BT;
  const int c0 = blockIdx.y * BCH;
  const int b = blockIdx.z;
  if (tid < 16) kf[tid] = __ldg(k4 + (3 - (tid >> 2)) * 4 + (3 - (tid & 3)));  // flipped

  const float* ub = u + (long long)b * hu * wu * ch;
  for (int idx = tid; idx < BIN * BIN * (BCH / 4); idx += 256) {
    const int c4 = idx & 7;
    const int pix = idx >> 3;
    const int r = pix / BIN, c = pix - r * BIN;
    const int iy = oy0 - 1 + r, ix = ox0 - 1 + c;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (iy >= 0 && iy < hu && ix >= 0 && ix < wu && c0 + c4 * 4 < ch)
      v = __ldg(reinterpret_cast<const float4*>(ub + ((long long)iy * wu + ix) * ch + c0 + c4 * 4));
    *reinterpret_cast<float4*>(&tile[pix * BCH + c4 * 4]) = v;
  }
  __syncthreads();

  const int c4 = tid & 7;
  const int p = tid >> 3;          // 0..31
  const int lx = p & 15;           // column inside the tile
  const int ly0 = (p >> 4) * 8;    // first of 8 rows
  const int cbase = c0 + c4 * 4;
  if (cbase >= ch) return;
  float kk[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) kk[i] = kf[i];
  float4 acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int r = 0; r < 11; ++r) {
    float4 row[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) row[e] = *reinterpret_cast<const float4*>(&tile[((ly0 + r) * BIN + lx + e) * BCH + c4 * 4]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int a = r - j;
      if (a >= 0 && a < 4) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float kv = kk[a * 4 + e];
          acc[j].x = fmaf(row[e].x, kv, acc[j].x);
          acc[j].y = fmaf(row[e].y, kv, acc[j].y);
          acc[j].z = fmaf(row[e].z, kv, acc[j].z);
          acc[j].w = fmaf(row[e].w, kv, acc[j].w);
        }
      }
    }
  }

  const int ox = ox0 + lx;
  if (ox >= ow) return;
  float4 bias = make_float4(0.f, 0.f, 0.f, 0.f), sn = make_float4(1.f, 1.f, 1.f, 1.f), dm = sn;
  float nw = 0.f;
  if (ep.d) dm = __ldg(reinterpret_cast<const float4*>(ep.d + (long long)b * ch + cbase));
  if (ep.activate) {
    if (ep.bias) bias = __ldg(reinterpret_cast<const float4*>(ep.bias + cbase));
    if (ep.noise) nw = __ldg(ep.noise_weight);
  }
  if (ep.s_next) sn = __ldg(reinterpret_cast<const float4*>(ep.s_next + (long long)b * ch + cbase));
  __nv_bfloat16* hi = reinterpret_cast<__nv_bfloat16*>(ep.out_hi);
  __nv_bfloat16* lo = reinterpret_cast<__nv_bfloat16*>(ep.out_lo);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int oy = oy0 + ly0 + j;
    if (oy >= oh) break;
    float v[4] = {__fmul_rn(acc[j].x, dm.x), __fmul_rn(acc[j].y, dm.y), __fmul_rn(acc[j].z, dm.z), __fmul_rn(acc[j].w, dm.w)};
    if (ep.activate) {
      float nz = 0.f;
      if (ep.noise) nz = __fmul_rn(nw, __ldg(ep.noise + (long long)b * ep.noise_bstride + (long long)oy * ow + ox));
      const float bb[4] = {bias.x, bias.y, bias.z, bias.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = lrelu_scaled(__fadd_rn(__fadd_rn(v[i], nz), bb[i]), ep.slope, ep.act_scale);
    }
    if (ep.out_f32_nchw) {
#pragma unroll
      for (int i = 0; i < 4; ++i) ep.out_f32_nchw[(((long long)b * ch + cbase + i) * oh + oy) * ow + ox] = v[i];
    }
    if (hi && ep.out_fmt == 1) {   // "f16" activation format: one fp16 plane
      const long long o = (((long long)b * oh + oy) * ow + ox) * ch + cbase;
      uint2 ph;
      ph.x = pack_f16x2_sat(v[0] * sn.x, v[1] * sn.y);
      ph.y = pack_f16x2_sat(v[2] * sn.z, v[3] * sn.w);
      *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(ep.out_hi) + o) = ph;
    } else if (hi) {
      const float ss[4] = {sn.x, sn.y, sn.z, sn.w};
      __nv_bfloat16 h4[4], l4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) split_bf16(v[i] * ss[i], h4[i], l4[i]);
      const long long o = (((long long)b * oh + oy) * ow + ox) * ch + cbase;
      *reinterpret_cast<uint2*>(hi + o) = *reinterpret_cast<const uint2*>(h4);
      *reinterpret_cast<uint2*>(lo + o) = *reinterpret_cast<const uint2*>(l4);
    }
  }
}


// ---------------------------------------------------------------------------------------------------------------
// TMA version (C % 32 == 0): persistent CTAs, the 19x19x32 fp32 input tile arrives by ONE cp.async.bulk.tensor per
// tile (zero padding = TMA out-of-bounds fill, no bounds logic), double-buffered through two mbarriers so that the
// load of tile i+2 overlaps the FMA/epilogue work of tile i+1.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_tile_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                                 int c3) {
  ptx::tma_load_4d(dst, m, bar, c0, c1, c2, c3);
}

// RPT = output rows per thread (8: 256 threads, 4: 512 threads — more warps to hide LDS / epilogue latency)
// FMT = what the epilogue writes, fixed at compile time for the two forms the generator uses (ncu, 1024^2 layer: 749 warp
//   instructions per tile and warp, ~90 of them the per-row format branches and the min/max pair of a runtime slope test):
//   0 = bf16 (hi, lo) planes only, 1 = one fp16 plane only, 2 = anything (fp32 NCHW map, no operand planes, ...).
// The leaky-ReLU is max(v, slope*v): the host routes slope > 1 to the generic kernel above.
template <int RPT, int FMT>
__global__ void __launch_bounds__(128 * (16 / RPT), 2) blur_act_nhwc_tma_kernel(const __grid_constant__ CUtensorMap tm_u,
                                                                   const float* __restrict__ k4, MauaConvEpilogue ep,
                                                                   int batch, int ch, int hu, int wu, int n_tiles,
                                                                   FastDiv fd_ncg, FastDiv fd_tx, FastDiv fd_ty) {
  using namespace ptx;
  constexpr uint32_t TILE_BYTES = BIN * BIN * BCH * 4;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
  // Byte offset of the 128-byte-aligned tile area inside the dynamic shared array.  The tile pointer is formed as
  // `smem_raw + offset` (pointer arithmetic on the __shared__ array) so that the compiler keeps the shared state space
  // and emits LDS.128: round-tripping through uintptr_t made it a GENERIC pointer kept in a 2-entry local-memory array
  // — every tile started with an LDL (19 % of the kernel's stall samples, ncu) followed by generic LD.E.128 loads.
  const uint32_t tile_off = base - smem_u32(smem_raw);
  const uint32_t bar0 = base + 2 * TILE_BYTES;
  __shared__ float kf[16];
  const int tid = threadIdx.x;
  const int oh = hu - 1, ow = wu - 1;
  const int tiles_x = (ow + BT - 1) / BT, tiles_y = (oh + BT - 1) / BT, ncg = ch / BCH;

  if (tid == 0) {
    prefetch_tmap(&tm_u);
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    fence_mbar_init();
  }
  pdl_launch_dependents();   // programmatic dependent launch (sm100_ptx.cuh): global memory only after the wait
  pdl_wait();
  if (tid < 16) kf[tid] = __ldg(k4 + (3 - (tid >> 2)) * 4 + (3 - (tid & 3)));  // flipped
  __syncthreads();

  // (multiply-shift divisions: the three runtime divisions were ~75 of the ~480 instructions per thread and tile)
  auto decode = [&](int t, int& cg, int& ox0, int& oy0, int& b) {
    const int r0 = fast_div(t, fd_ncg);
    cg = t - r0 * ncg;
    const int r1 = fast_div(r0, fd_tx);
    ox0 = (r0 - r1 * tiles_x) * BT;
    b = fast_div(r1, fd_ty);
    oy0 = (r1 - b * tiles_y) * BT;
  };
  auto issue = [&](int t, int stage) {
    int cg, ox0, oy0, b;
    decode(t, cg, ox0, oy0, b);
    mbar_expect_tx(bar0 + 8 * stage, TILE_BYTES);
    tma_load_tile_4d(base + stage * TILE_BYTES, &tm_u, bar0 + 8 * stage, cg * BCH, ox0 - 1, oy0 - 1, b);
  };

  const int first = blockIdx.x, step = gridDim.x;
  if (tid == 0) {
    if (first < n_tiles) issue(first, 0);
    if (first + step < n_tiles) issue(first + step, 1);
  }
  float kk[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) kk[i] = kf[i];
  // Separable FIR (the generator's kernels are outer products: make_kernel([1,3,3,1]), models/stylegan2.py:23-31):
  // K[a][e] = ka[a] * kb[e] evaluated as a horizontal 4-tap pass per staged row followed by the vertical taps — 152 instead of
  // 256 FFMA2 per thread and tile (the FMA pipe was the busiest unit of this kernel: 43 % with 55 % of the issue slots
  // taken, ncu).  Exact factorisation for dyadic taps; any other 4x4 kernel takes the general 16-tap path below.
  float ka[4], kb[4];
  bool sep = kk[0] != 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    kb[i] = kk[i];
    ka[i] = sep ? kk[4 * i] / kk[0] : 0.f;
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int e = 0; e < 4; ++e) sep = sep && (ka[a] * kb[e] == kk[4 * a + e]);
  const int c4 = tid & 7;
  const int p = tid >> 3;
  const int lx = p & 15;
  const int ly0 = (p >> 4) * RPT;
  const bool hi = ep.out_hi != nullptr;
  const float nw = (ep.activate && ep.noise) ? __ldg(ep.noise_weight) : 0.f;

  int it = 0;
  for (int t = first; t < n_tiles; t += step, ++it) {
    const int stage = it & 1;
    // tile coordinates and epilogue operands first: these global loads overlap the TMA wait and the FMA phase
    int cg, ox0, oy0, b;
    decode(t, cg, ox0, oy0, b);
    const int cbase = cg * BCH + c4 * 4;
    const int ox = ox0 + lx;
    float4 bias = make_float4(0.f, 0.f, 0.f, 0.f), sn = make_float4(1.f, 1.f, 1.f, 1.f), dm = sn;
    float nzv[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) nzv[j] = 0.f;
    if (ox < ow) {
      if (ep.activate && ep.bias) bias = __ldg(reinterpret_cast<const float4*>(ep.bias + cbase));
      if (ep.s_next) sn = __ldg(reinterpret_cast<const float4*>(ep.s_next + (long long)b * ch + cbase));
      if (ep.d) dm = __ldg(reinterpret_cast<const float4*>(ep.d + (long long)b * ch + cbase));
      if (ep.activate && ep.noise) {
        const float* np = ep.noise + (long long)b * ep.noise_bstride + (long long)(oy0 + ly0) * ow + ox;
#pragma unroll
        for (int j = 0; j < RPT; ++j)
          if (oy0 + ly0 + j < oh) nzv[j] = __ldg(np + (long long)j * ow);
      }
    }
    mbar_wait(bar0 + 8 * stage, (it >> 1) & 1);
    const float* tile = reinterpret_cast<const float*>(smem_raw + tile_off + (uint32_t)stage * TILE_BYTES);
    float4 acc[RPT];
#pragma unroll
    for (int j = 0; j < RPT; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (sep) {
#pragma unroll
      for (int r = 0; r < RPT + 3; ++r) {
        float2 hl = make_float2(0.f, 0.f), hh = hl;   // horizontal pass of staged row r (4 channels)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float4 v = *reinterpret_cast<const float4*>(&tile[((ly0 + r) * BIN + lx + e) * BCH + c4 * 4]);
          const float2 kv = make_float2(kb[e], kb[e]);
          hl = ffma2(make_float2(v.x, v.y), kv, hl);
          hh = ffma2(make_float2(v.z, v.w), kv, hh);
        }
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
          const int a = r - j;
          if (a >= 0 && a < 4) {
            const float2 kv = make_float2(ka[a], ka[a]);
            const float2 lo2 = ffma2(hl, kv, make_float2(acc[j].x, acc[j].y));
            const float2 hi2 = ffma2(hh, kv, make_float2(acc[j].z, acc[j].w));
            acc[j] = make_float4(lo2.x, lo2.y, hi2.x, hi2.y);
          }
        }
      }
    } else {
#pragma unroll
      for (int r = 0; r < RPT + 3; ++r) {
        float4 row[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          row[e] = *reinterpret_cast<const float4*>(&tile[((ly0 + r) * BIN + lx + e) * BCH + c4 * 4]);
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
          const int a = r - j;
          if (a >= 0 && a < 4) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 kv = make_float2(kk[a * 4 + e], kk[a * 4 + e]);
              const float2 lo2 = ffma2(make_float2(row[e].x, row[e].y), kv, make_float2(acc[j].x, acc[j].y));
              const float2 hi2 = ffma2(make_float2(row[e].z, row[e].w), kv, make_float2(acc[j].z, acc[j].w));
              acc[j] = make_float4(lo2.x, lo2.y, hi2.x, hi2.y);
            }
          }
        }
      }
    }
    __syncthreads();  // every thread has consumed this stage: refill it while the epilogue runs
    if (tid == 0 && t + 2 * step < n_tiles) {
      fence_proxy_async();
      issue(t + 2 * step, stage);
    }

    if (ox >= ow) continue;
    // Epilogue.  ncu (1024^2 layer): this kernel is issue-bound, not memory-bound — 968 warp instructions per tile and warp,
    // of which only 152 FFMA2 + 44 LDS are the FIR; the rest was scalar activation math (136 FMUL, 64 FADD, 32 FSETP/FSEL
    // pairs) and 64-bit address arithmetic per row.  Now: packed f32x2 math with the activation gain folded into s_next,
    // leaky-ReLU as max(v, slope*v) (min for slope > 1), and per-tile base pointers advanced by a constant row stride.
    const bool act = ep.activate != 0;
    const float slope = act ? ep.slope : 1.f, gain = act ? ep.act_scale : 1.f;
    const float2 sl2 = make_float2(slope, slope);
    const float2 d_lo = make_float2(dm.x, dm.y), d_hi = make_float2(dm.z, dm.w);
    const float2 b_lo = make_float2(bias.x, bias.y), b_hi = make_float2(bias.z, bias.w);
    const float2 s_lo = make_float2(sn.x * gain, sn.y * gain), s_hi = make_float2(sn.z * gain, sn.w * gain);
    const int oy_first = oy0 + ly0;
    const long long pix0 = ((long long)b * oh + oy_first) * ow + ox;
    const long long row_elems = (long long)ow * ch;
    uint16_t* ph = (FMT != 2 || hi) ? reinterpret_cast<uint16_t*>(ep.out_hi) + pix0 * ch + cbase : nullptr;
    uint16_t* pl = (FMT == 0 || (FMT == 2 && hi && ep.out_fmt == 0)) ? reinterpret_cast<uint16_t*>(ep.out_lo) + pix0 * ch + cbase
                                                                      : nullptr;
    float* pn = (FMT == 2 && ep.out_f32_nchw) ? ep.out_f32_nchw + (((long long)b * ch + cbase) * oh + oy_first) * ow + ox : nullptr;
    const long long plane = (long long)oh * ow;
    const bool f16out = FMT == 1 || (FMT == 2 && ep.out_fmt == 1);
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      if (oy_first + j >= oh) break;
      const float nz = __fmul_rn(nw, nzv[j]);
      const float2 nz2 = make_float2(nz, nz);
      // v = FIR * d + (noise + bias)   [d: demodulation left to this kernel by the transposed conv, commutes with the FIR]
      float2 v_lo = ffma2(make_float2(acc[j].x, acc[j].y), d_lo, fadd2(b_lo, nz2));
      float2 v_hi = ffma2(make_float2(acc[j].z, acc[j].w), d_hi, fadd2(b_hi, nz2));
      const float2 t_lo = fmul2(v_lo, sl2), t_hi = fmul2(v_hi, sl2);
      v_lo = make_float2(fmaxf(v_lo.x, t_lo.x), fmaxf(v_lo.y, t_lo.y));
      v_hi = make_float2(fmaxf(v_hi.x, t_hi.x), fmaxf(v_hi.y, t_hi.y));
      if (FMT == 2 && pn) {
        float* o = pn + (long long)j * ow;
        o[0] = v_lo.x * gain; o[plane] = v_lo.y * gain; o[2 * plane] = v_hi.x * gain; o[3 * plane] = v_hi.y * gain;
      }
      if (FMT != 2 || ph) {
        const float2 a_lo = fmul2(v_lo, s_lo), a_hi = fmul2(v_hi, s_hi);     // * gain * s_next
        uint16_t* dst = ph + (long long)j * row_elems;
        if (f16out) {   // "f16" activation format: one fp16 plane (8-byte stores)
          uint2 o;
          o.x = pack_f16x2_sat(a_lo.x, a_lo.y);
          o.y = pack_f16x2_sat(a_hi.x, a_hi.y);
          *reinterpret_cast<uint2*>(dst) = o;
        } else {
          const __nv_bfloat162 h01 = __floats2bfloat162_rn(a_lo.x, a_lo.y), h23 = __floats2bfloat162_rn(a_hi.x, a_hi.y);
          const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
          const float2 neg1 = make_float2(-1.f, -1.f);
          const float2 r01 = ffma2(f01, neg1, a_lo), r23 = ffma2(f23, neg1, a_hi);   // exact: a - hi
          const __nv_bfloat162 l01 = __floats2bfloat162_rn(r01.x, r01.y), l23 = __floats2bfloat162_rn(r23.x, r23.y);
          uint2 oh2, ol2;
          oh2.x = *reinterpret_cast<const uint32_t*>(&h01); oh2.y = *reinterpret_cast<const uint32_t*>(&h23);
          ol2.x = *reinterpret_cast<const uint32_t*>(&l01); ol2.y = *reinterpret_cast<const uint32_t*>(&l23);
          *reinterpret_cast<uint2*>(dst) = oh2;
          *reinterpret_cast<uint2*>(pl + (long long)j * row_elems) = ol2;
        }
      }
    }
  }
}

}  // namespace maua

extern "C" int maua_blur_act_nhwc(const float* u, const float* k4, const MauaConvEpilogue* ep_host, int batch, int ch,
                                  int hu, int wu, void* stream) {
  using namespace maua;
  MAUA_CHECK_ARG(u && k4 && ep_host && batch >= 0 && ch >= 4 && hu >= 2 && wu >= 2, "blur_act_nhwc: bad arguments");
  MAUA_CHECK_ARG(ch % 4 == 0, "blur_act_nhwc: channels must be a multiple of 4");
  MAUA_CHECK_ARG(ep_host->out_fmt == 0 || ep_host->out_fmt == 1, "blur_act_nhwc: out_fmt must be 0 or 1");
  MAUA_CHECK_ARG(ep_host->out_fmt == 1 || (ep_host->out_hi != nullptr) == (ep_host->out_lo != nullptr),
                 "blur_act_nhwc: hi/lo must come in pairs");
  MAUA_CHECK_ARG(!ep_host->noise || ep_host->noise_weight, "blur_act_nhwc: noise without weight");
  if (batch == 0) return MAUA_OK;
  MAUA_CHECK_ARG(batch <= 65535, "blur_act_nhwc: batch too large");
  const int oh = hu - 1, ow = wu - 1;
  const bool slope_ok = !ep_host->activate || ep_host->slope <= 1.f;   // lrelu as max(v, slope*v)
  if (ch % BCH == 0 && (reinterpret_cast<uintptr_t>(u) & 15) == 0 && slope_ok) {
    CUtensorMap tm;
    const cuuint64_t dims[4] = {(cuuint64_t)ch, (cuuint64_t)wu, (cuuint64_t)hu, (cuuint64_t)batch};
    const cuuint64_t strides[3] = {(cuuint64_t)ch * 4, (cuuint64_t)wu * ch * 4, (cuuint64_t)hu * wu * ch * 4};
    const cuuint32_t box[4] = {BCH, BIN, BIN, 1};
    int rc = tmap::encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, u, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc != MAUA_OK) return rc;
    const long long n_tiles = (long long)ceil_div(ow, BT) * ceil_div(oh, BT) * (ch / BCH) * batch;
    MAUA_CHECK_ARG(n_tiles < (1LL << 31), "blur_act_nhwc: too many tiles");
    const size_t smem = 2 * (size_t)BIN * BIN * BCH * 4 + 16 + 128;
    static const int rpt = [] { const char* e = getenv("MAUA_BLUR_RPT"); return e ? atoi(e) : 8; }();
    const int n_sm = device_sm_count();
    const int grid = (int)(n_tiles < n_sm * 2 ? n_tiles : n_sm * 2);
    const FastDiv f0 = make_fastdiv((uint32_t)(ch / BCH)), f1 = make_fastdiv((uint32_t)ceil_div(ow, BT)),
                  f2 = make_fastdiv((uint32_t)ceil_div(oh, BT));
    // operand planes only (what every transposed layer of the generator that feeds a tensor-core layer writes) -> the
    // compile-time formats; anything else (fp32 NCHW map for ToRGB / bends, no planes) -> the generic epilogue
    const int fmt = (ep_host->out_f32_nchw || !ep_host->out_hi || rpt != 8) ? 2 : (ep_host->out_fmt == 1 ? 1 : 0);
#define MAUA_BLUR_LAUNCH(RPTV, FMTV)                                                                                  \
  do {                                                                                                                \
    MAUA_CHECK_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(blur_act_nhwc_tma_kernel<RPTV, FMTV>), smem));      \
    MAUA_CHECK_CUDA(launch_chain(blur_act_nhwc_tma_kernel<RPTV, FMTV>, dim3(grid), dim3(128 * (16 / RPTV)), smem,     \
                                 as_stream(stream), 1, tm, k4, *ep_host, batch, ch, hu, wu, (int)n_tiles, f0, f1, f2)); \
  } while (0)
    if (rpt != 8) MAUA_BLUR_LAUNCH(4, 2);
    else if (fmt == 0) MAUA_BLUR_LAUNCH(8, 0);
    else if (fmt == 1) MAUA_BLUR_LAUNCH(8, 1);
    else MAUA_BLUR_LAUNCH(8, 2);
#undef MAUA_BLUR_LAUNCH
    MAUA_CHECK_LAUNCH("blur_act_nhwc(tma)");
    return MAUA_OK;
  }
  dim3 grid(ceil_div(ow, BT) * ceil_div(oh, BT), ceil_div(ch, BCH), batch);
  blur_act_nhwc_kernel<<<grid, 256, 0, as_stream(stream)>>>(u, k4, *ep_host, ch, hu, wu);
  MAUA_CHECK_LAUNCH("blur_act_nhwc");
  return MAUA_OK;
}
