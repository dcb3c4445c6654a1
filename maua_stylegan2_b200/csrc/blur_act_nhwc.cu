// Blur (4x4 FIR, pad (1,1)) + noise + bias + leaky-ReLU + next-layer style + bf16 split, channels-last.
//
// Second half of an upsampling StyledConv on the tensor-core path (models/stylegan2.py:238 Blur.forward,
// :262-266 NoiseInjection, op/fused_act.py:82-97): the transposed-conv kernel leaves u = d * convT(x) as fp32
// [B, 2H+1, 2W+1, C]; this kernel computes
//     o[b,oy,ox,c] = sum_{a,e<4} Kf[a][e] * u[b, oy+a-1, ox+e-1, c]        (Kf = flipped FIR, zero outside u)
//     o = lrelu(o + nw*noise[b,oy,ox] + bias[c]) * sqrt2
// and writes what the next layer consumes: (hi, lo) bf16 NHWC of o * s_next[b,c] and/or fp32 NCHW o.
// HBM-bound: 4 B/elt in, 4 B/elt out.  One CTA = 16x16 output pixels x 32 channels; the 19x19x32 input tile
// is staged in shared memory with 128-byte-contiguous loads; each thread slides down a column of 8 outputs for
// one float4 of channels (44 LDS.128 per 8 outputs), stores are 8-byte bf16x4 per plane.
#include "common.cuh"

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
  float4 bias = make_float4(0.f, 0.f, 0.f, 0.f), sn = make_float4(1.f, 1.f, 1.f, 1.f);
  float nw = 0.f;
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
    float v[4] = {acc[j].x, acc[j].y, acc[j].z, acc[j].w};
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
    if (hi) {
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

}  // namespace maua

extern "C" int maua_blur_act_nhwc(const float* u, const float* k4, const MauaConvEpilogue* ep_host, int batch, int ch,
                                  int hu, int wu, void* stream) {
  using namespace maua;
  MAUA_CHECK_ARG(u && k4 && ep_host && batch >= 0 && ch >= 4 && hu >= 2 && wu >= 2, "blur_act_nhwc: bad arguments");
  MAUA_CHECK_ARG(ch % 4 == 0, "blur_act_nhwc: channels must be a multiple of 4");
  MAUA_CHECK_ARG((ep_host->out_hi != nullptr) == (ep_host->out_lo != nullptr), "blur_act_nhwc: hi/lo must come in pairs");
  MAUA_CHECK_ARG(!ep_host->noise || ep_host->noise_weight, "blur_act_nhwc: noise without weight");
  if (batch == 0) return MAUA_OK;
  MAUA_CHECK_ARG(batch <= 65535, "blur_act_nhwc: batch too large");
  const int oh = hu - 1, ow = wu - 1;
  dim3 grid(ceil_div(ow, BT) * ceil_div(oh, BT), ceil_div(ch, BCH), batch);
  blur_act_nhwc_kernel<<<grid, 256, 0, as_stream(stream)>>>(u, k4, *ep_host, ch, hu, wu);
  MAUA_CHECK_LAUNCH("blur_act_nhwc");
  return MAUA_OK;
}
