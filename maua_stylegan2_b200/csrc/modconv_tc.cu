// ModulatedConv2d (3x3) on the 5th-generation tensor cores: TMA -> shared memory -> tcgen05.mma -> TMEM -> epilogue.
//
// Reference path replaced: models/stylegan2.py:217-254 (per-sample weight modulation + cuDNN grouped conv /
// grouped transposed conv) + :262-266 (noise) + op/fused_act.py:82-97 (bias + leaky ReLU).
//
// Formulation:  o[b,co,p] = d[b,co] * sum_{tap,ci} Wc[tap][co][ci] * (s[b,ci]*x[b,ci,p+tap])
//   - the style s is already folded into the activation operand by the producing kernel, so the weight operand
//     is shared by the whole batch and only d (per sample, per output channel) remains for the epilogue;
//   - implicit GEMM: M = 128 output pixels (TB x TH x TW block of [batch, y, x]), N = BN output channels,
//     K = Cin per tap.  For tap (ky,kx) the A tile is ONE 4-D TMA box {KC ch, TW, TH, TB} of the NHWC activation
//     shifted by (dy,dx); the zero padding of the convolution is the TMA out-of-bounds fill, so there is no
//     im2col buffer and no halo logic.  B tile = {KC, BN, 1} box of the packed [tap][Cout][Cin] weight.
//   - precision: operands are (hi, lo) bf16 pairs; the tensor core evaluates hi*hi + hi*lo + lo*hi into one fp32
//     TMEM accumulator (n_products = 3) => ~2^-16 relative product error, fp32-grade results on the bf16 pipe.
//   - transposed (up) layers are evaluated as 4 sub-pixel phases with 4 TMEM accumulators; only 4 distinct
//     shifted A tiles serve the 9 taps (A ring stage is held across the taps that share a shift).
//
// Warp roles (192 threads): warp 0 = TMA producer (1 lane), warp 1 = TMEM owner + MMA issuer (1 lane),
// warps 2-5 = epilogue (tcgen05.ld 32 lanes x 16 columns, demod/noise/bias/lrelu/next-style/split, stores).
// Two independent mbarrier rings (A tiles, B tiles) + one accumulator-full barrier.
#include <cuda.h>
#include <cstdlib>

#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tmap.cuh"

namespace maua {
namespace tc {

struct Step {
  int8_t dy, dx;    // shift of the A tile
  int8_t a_new;     // 1: this step loads / consumes a new A stage
  int8_t a_last;    // 1: last step that reads the current A stage
  int8_t tap;       // ky*3 + kx  (index into the packed weight)
  int8_t phase;     // TMEM accumulator index (sub-pixel phase for up layers)
  int8_t pad0, pad1;
};

// same-resolution: tap (ky,kx) reads x[y+ky-1, x+kx-1]
// up (stride-2 transposed): u[2y'+py, 2x'+px] += W[ky,kx] * x[y'+dy, x'+dx], py = ky&1, dy = (ky==2 ? -1 : 0)
__constant__ Step c_steps[2][9] = {
    {{-1, -1, 1, 1, 0, 0}, {-1, 0, 1, 1, 1, 0}, {-1, 1, 1, 1, 2, 0}, {0, -1, 1, 1, 3, 0}, {0, 0, 1, 1, 4, 0},
     {0, 1, 1, 1, 5, 0},   {1, -1, 1, 1, 6, 0}, {1, 0, 1, 1, 7, 0},  {1, 1, 1, 1, 8, 0}},
    {{0, 0, 1, 0, 0, 0},   {0, 0, 0, 0, 1, 1},  {0, 0, 0, 0, 3, 2},  {0, 0, 0, 1, 4, 3},  {0, -1, 1, 0, 2, 0},
     {0, -1, 0, 1, 5, 2},  {-1, 0, 1, 0, 6, 0}, {-1, 0, 0, 1, 7, 1}, {-1, -1, 1, 1, 8, 0}}};

struct Params {
  int B, H, W, Cin, Cout;  // input activation dims / channels
  int GH, GW;              // GEMM pixel grid: (H, W) or (H+1, W+1) for up
  int TB, TH, TW;          // M tile = TB*TH*TW <= 128 rows (any factors: 17x7 covers the 17x17 grid of a 16^2 up layer)
  int rows;                // TB*TH*TW: rows of the 128-row A tile the TMA box fills (the rest is never read back)
  int S;                   // split-K: S CTAs share one output tile, each reduces n_kchunks/S K chunks
  int tiles_x, tiles_y;
  int BN, n_tiles;
  int n_kchunks;
  int SA, SB;
  int nprod;
  uint32_t tmem_cols;
};

__device__ __forceinline__ float lrelu_s(float v, float slope, float scale) {
  return (v > 0.f ? v : v * slope) * scale;
}

template <int KC, bool UP>
__global__ void __launch_bounds__(192, 1)
modconv_tc_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                  const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                  const Params p, const MauaConvEpilogue ep) {
  using namespace ptx;
  constexpr uint32_t ROW_BYTES = KC * 2;               // one swizzle span per operand row
  constexpr uint32_t A_HALF = 128 * ROW_BYTES;         // one bf16 plane of the A tile
  constexpr uint32_t A_STAGE = 2 * A_HALF;             // hi + lo
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024-byte alignment
  const uint32_t b_half = (uint32_t)p.BN * ROW_BYTES;
  const uint32_t b_stage = 2 * b_half;
  const uint32_t a_base = smem0;
  const uint32_t b_base = a_base + (uint32_t)p.SA * A_STAGE;
  const uint32_t bar_base = b_base + (uint32_t)p.SB * b_stage;
  // barrier slots (8 B each): a_full[SA], a_empty[SA], b_full[SB], b_empty[SB], acc_full, tmem_ptr
  const uint32_t a_full = bar_base, a_empty = a_full + 8 * p.SA;
  const uint32_t b_full = a_empty + 8 * p.SA, b_empty = b_full + 8 * p.SB;
  const uint32_t acc_full = b_empty + 8 * p.SB;
  const uint32_t tmem_slot = acc_full + 8;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- tile coordinates ------------------------------------------------------------------------------------------
  const int split = blockIdx.x % p.S;
  const int tile_id = blockIdx.x / p.S;
  const int n_tile = tile_id % p.n_tiles;
  const int m_tile = tile_id / p.n_tiles;
  const int kc_per = p.n_kchunks / p.S, kc_begin = split * kc_per, kc_end = kc_begin + kc_per;
  const int tile_x = m_tile % p.tiles_x;
  const int tile_y = (m_tile / p.tiles_x) % p.tiles_y;
  const int tile_b = m_tile / (p.tiles_x * p.tiles_y);
  const int x0 = tile_x * p.TW, y0 = tile_y * p.TH, b0 = tile_b * p.TB, n0 = n_tile * p.BN;

  // ---- one-time setup --------------------------------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_a_hi);
    prefetch_tmap(&tm_b_hi);
    if (p.nprod > 1) {
      prefetch_tmap(&tm_a_lo);
      prefetch_tmap(&tm_b_lo);
    }
    for (int i = 0; i < p.SA; ++i) {
      mbar_init(a_full + 8 * i, 1);
      mbar_init(a_empty + 8 * i, 1);
    }
    for (int i = 0; i < p.SB; ++i) {
      mbar_init(b_full + 8 * i, 1);
      mbar_init(b_empty + 8 * i, 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_launch_dependents();   // programmatic dependent launch: nothing above touched global memory (sm100_ptx.cuh)
  pdl_wait();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const Step* steps = c_steps[UP ? 1 : 0];
  const uint32_t a_bytes = (uint32_t)p.rows * ROW_BYTES * (p.nprod > 1 ? 2u : 1u);  // what the TMA boxes deliver
  const uint32_t b_bytes = (p.nprod > 1) ? b_stage : b_half;

  if (warp == 0 && lane == 0) {
    // ================================ TMA producer ================================
    int ia = 0, ib = 0;
    uint32_t pa = 0, pb = 0;
    for (int kc = kc_begin; kc < kc_end; ++kc) {
      const int c0 = kc * KC;
#pragma unroll 1
      for (int st = 0; st < 9; ++st) {
        const Step s = steps[st];
        if (s.a_new) {
          mbar_wait(a_empty + 8 * ia, pa ^ 1);
          mbar_expect_tx(a_full + 8 * ia, a_bytes);
          const uint32_t dst = a_base + ia * A_STAGE;
          tma_load_4d(dst, &tm_a_hi, a_full + 8 * ia, c0, x0 + s.dx, y0 + s.dy, b0);
          if (p.nprod > 1) tma_load_4d(dst + A_HALF, &tm_a_lo, a_full + 8 * ia, c0, x0 + s.dx, y0 + s.dy, b0);
          if (++ia == p.SA) { ia = 0; pa ^= 1; }
        }
        mbar_wait(b_empty + 8 * ib, pb ^ 1);
        mbar_expect_tx(b_full + 8 * ib, b_bytes);
        const uint32_t dstb = b_base + ib * b_stage;
        tma_load_3d(dstb, &tm_b_hi, b_full + 8 * ib, c0, n0, s.tap);
        if (p.nprod > 1) tma_load_3d(dstb + b_half, &tm_b_lo, b_full + 8 * ib, c0, n0, s.tap);
        if (++ib == p.SB) { ib = 0; pb ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // whole warp, uniform control flow; one elected lane issues (keeps descriptors in uniform registers)
    const bool leader = elect_one_sync();
    const uint32_t idesc = make_idesc_bf16(128, (uint32_t)p.BN);
    const uint64_t da0 = make_kmajor_desc(a_base, ROW_BYTES), db0 = make_kmajor_desc(b_base, ROW_BYTES);
    const uint64_t a_stage16 = A_STAGE >> 4, a_half16 = A_HALF >> 4, b_stage16 = b_stage >> 4, b_half16 = b_half >> 4;
    int ia = 0, ib = 0, cur_a = 0;
    uint32_t pa = 0, pb = 0, started = 0;
    for (int kc = kc_begin; kc < kc_end; ++kc) {
#pragma unroll 1
      for (int st = 0; st < 9; ++st) {
        const Step s = steps[st];
        if (s.a_new) {
          cur_a = ia;
          mbar_wait(a_full + 8 * ia, pa);
          if (++ia == p.SA) { ia = 0; pa ^= 1; }
        }
        mbar_wait(b_full + 8 * ib, pb);
        tc_fence_after();
        const uint64_t dah = da0 + (uint64_t)cur_a * a_stage16, dal = dah + a_half16;
        const uint64_t dbh = db0 + (uint64_t)ib * b_stage16, dbl = dbh + b_half16;
        const uint32_t acc = tmem_base + (uint32_t)s.phase * (uint32_t)p.BN;
        const uint32_t first = (started >> s.phase) & 1u;
        if (leader) {
#pragma unroll
          for (int k = 0; k < KC / 16; ++k) umma_bf16(acc, dah + 2 * k, dbh + 2 * k, idesc, k == 0 ? first : 1u);
          if (p.nprod > 1) {
#pragma unroll
            for (int k = 0; k < KC / 16; ++k) umma_bf16(acc, dah + 2 * k, dbl + 2 * k, idesc, 1u);
#pragma unroll
            for (int k = 0; k < KC / 16; ++k) umma_bf16(acc, dal + 2 * k, dbh + 2 * k, idesc, 1u);
          }
          umma_commit(b_empty + 8 * ib);
          if (s.a_last) umma_commit(a_empty + 8 * cur_a);
        }
        started |= 1u << s.phase;
        if (++ib == p.SB) { ib = 0; pb ^= 1; }
      }
    }
    if (leader) umma_commit(acc_full);
  } else if (warp >= 2) {
    // ================================ epilogue ================================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int m = quad * 32 + lane;
    const int tx = m % p.TW, ty = (m / p.TW) % p.TH, tb = m / (p.TW * p.TH);
    const int gx = x0 + tx, gy = y0 + ty, b = b0 + tb;
    const bool in_grid = (m < p.rows) && (b < p.B) && (gy < p.GH) && (gx < p.GW);
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
    constexpr int NPH = UP ? 4 : 1;
    // ---- deterministic split-K: every CTA parks its partial accumulators in the workspace; the CTA that arrives last
    // at the tile's counter sums the S partials in split order (so the result does not depend on which CTA that is)
    // and runs the epilogue.  Layout: [4 KB counters][tile][split][16-column chunk][row m][16 floats] (coalesced).
    const int n_chunks = NPH * p.BN / 16;
    float* ws_tile = nullptr;
    if (p.S > 1) {
      ws_tile = reinterpret_cast<float*>(static_cast<char*>(ep.workspace) + 4096) +
                (size_t)tile_id * p.S * n_chunks * 128 * 16;
      float* mine = ws_tile + ((size_t)split * n_chunks * 128 + m) * 16;
#pragma unroll 1
      for (int ch = 0; ch < n_chunks; ++ch) {
        uint32_t r[16];
        tmem_ld_x16(lane_addr + (uint32_t)(ch * 16), r);
        tmem_ld_wait();
        float4* dst = reinterpret_cast<float4*>(mine + (size_t)ch * 128 * 16);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          dst[i] = make_float4(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1]), __uint_as_float(r[4 * i + 2]),
                               __uint_as_float(r[4 * i + 3]));
      }
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      __shared__ int s_last;
      if (threadIdx.x == 64) {
        unsigned* counter = static_cast<unsigned*>(ep.workspace) + tile_id;
        const unsigned old = atomicAdd(counter, 1u);
        s_last = (old == (unsigned)p.S - 1u);
        if (s_last) *counter = 0u;  // workspace is handed back zeroed (the next launch is stream-ordered)
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (!s_last) ws_tile = nullptr;
      else __threadfence();
    }
    const bool do_epilogue = (p.S == 1) || (ws_tile != nullptr);
    const int bc = in_grid ? b : 0;
    const float* dptr = ep.d ? ep.d + (long long)bc * p.Cout + n0 : nullptr;
#pragma unroll 1
    for (int ph = 0; ph < (do_epilogue ? NPH : 0); ++ph) {
      int oy, ox, OH, OW;
      bool valid;
      if (UP) {
        OH = 2 * p.H + 1; OW = 2 * p.W + 1;
        oy = 2 * gy + (ph >> 1); ox = 2 * gx + (ph & 1);
        valid = in_grid && oy < OH && ox < OW;
      } else {
        OH = p.H; OW = p.W; oy = gy; ox = gx; valid = in_grid;
      }
      const long long pix = valid ? (((long long)b * OH + oy) * OW + ox) : 0;
      float nz = 0.f;
      if (!UP && ep.activate && ep.noise && valid)
        nz = __ldg(ep.noise_weight) * __ldg(ep.noise + (long long)b * ep.noise_bstride + (long long)oy * OW + ox);
#pragma unroll 1
      for (int c = 0; c < p.BN; c += 16) {
        float v[16];
        if (ws_tile) {
          if (!valid) continue;
          const float* src = ws_tile + ((size_t)((ph * p.BN + c) >> 4) * 128 + m) * 16;
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = 0.f;
          const size_t sp_stride = (size_t)n_chunks * 128 * 16;
#pragma unroll 1
          for (int sp0 = 0; sp0 < p.S; sp0 += 4) {  // 4 partials in flight (L2 latency), summed in split order
            float4 t[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (sp0 + j < p.S) {
                const float4* q = reinterpret_cast<const float4*>(src + (size_t)(sp0 + j) * sp_stride);
#pragma unroll
                for (int i = 0; i < 4; ++i) t[j][i] = __ldcg(q + i);
              }
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (sp0 + j < p.S) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  v[4 * i] += t[j][i].x; v[4 * i + 1] += t[j][i].y; v[4 * i + 2] += t[j][i].z; v[4 * i + 3] += t[j][i].w;
                }
              }
          }
        } else {
          uint32_t r[16];
          tmem_ld_x16(lane_addr + (uint32_t)(ph * p.BN + c), r);
          tmem_ld_wait();
          if (!valid) continue;
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
        }
        if (dptr) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] *= __ldg(dptr + c + i);
        }
        if (UP) {
          float4* dst = reinterpret_cast<float4*>(ep.out_raw_nhwc + pix * p.Cout + n0 + c);
#pragma unroll
          for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
          if (ep.activate) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float bb = ep.bias ? __ldg(ep.bias + n0 + c + i) : 0.f;
              v[i] = lrelu_s((v[i] + nz) + bb, ep.slope, ep.act_scale);
            }
          }
          if (ep.out_f32_nchw) {
            float* dst = ep.out_f32_nchw + (((long long)b * p.Cout + n0 + c) * OH + oy) * OW + ox;
            const long long plane = (long long)OH * OW;
#pragma unroll
            for (int i = 0; i < 16; ++i) dst[i * plane] = v[i];
          }
          if (ep.out_hi) {
            __nv_bfloat16 h[16], l[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float sv = ep.s_next ? __ldg(ep.s_next + (long long)b * p.Cout + n0 + c + i) : 1.f;
              split_bf16(v[i] * sv, h[i], l[i]);
            }
            uint4* dh = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out_hi) + pix * p.Cout + n0 + c);
            uint4* dl = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(ep.out_lo) + pix * p.Cout + n0 + c);
            dh[0] = reinterpret_cast<const uint4*>(h)[0];
            dh[1] = reinterpret_cast<const uint4*>(h)[1];
            dl[0] = reinterpret_cast<const uint4*>(l)[0];
            dl[1] = reinterpret_cast<const uint4*>(l)[1];
          }
        }
      }
    }
  }

  // ---- teardown --------------------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
static int encode_bf16(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                       const cuuint32_t* box, int row_bytes) {
  return tmap::encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_b, box,
                      row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}

}  // namespace tc
}  // namespace maua

namespace maua {
int modconv_tc2_launch(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo,
                       const MauaConvEpilogue& ep, int batch, int cin, int cout, int h, int w, int up, int n_products,
                       cudaStream_t st);
}

extern "C" int maua_modconv_tc(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo,
                               const MauaConvEpilogue* ep_host, int batch, int cin, int cout, int h, int w, int up,
                               int n_products, void* stream) {
  using namespace maua;
  using namespace maua::tc;
  MAUA_CHECK_ARG(x_hi && w_hi && ep_host, "modconv_tc: null pointer");
  MAUA_CHECK_ARG(n_products >= 1 && n_products <= 3, "modconv_tc: n_products must be 1, 2 (fp16) or 3");
  MAUA_CHECK_ARG(n_products != 3 || (x_lo && w_lo), "modconv_tc: lo planes required for n_products == 3");
  MAUA_CHECK_ARG(n_products != 2 || w_lo, "modconv_tc: the fp16 mode needs the fp16 (hi, lo) weight pair");
  MAUA_CHECK_ARG(batch >= 0 && h >= 1 && w >= 1, "modconv_tc: bad shape");
  MAUA_CHECK_ARG(cin % 32 == 0 && cin >= 32, "modconv_tc: Cin must be a multiple of 32");
  MAUA_CHECK_ARG(cout % 16 == 0 && cout >= 16, "modconv_tc: Cout must be a multiple of 16");
  const MauaConvEpilogue& ep = *ep_host;
  if (up) {
    MAUA_CHECK_ARG(ep.out_raw_nhwc, "modconv_tc(up): out_raw_nhwc required");
  } else {
    MAUA_CHECK_ARG(ep.out_f32_nchw || ep.out_hi || ep.rgb_out, "modconv_tc: no output requested");
    MAUA_CHECK_ARG((ep.rgb_out != nullptr) == (ep.rgb_w != nullptr), "modconv_tc: rgb_w / rgb_out must come in pairs");
    MAUA_CHECK_ARG(ep.out_fmt == 0 || ep.out_fmt == 1, "modconv_tc: out_fmt must be 0 (bf16 pair) or 1 (fp16 plane)");
    MAUA_CHECK_ARG(ep.out_fmt == 1 || (ep.out_hi != nullptr) == (ep.out_lo != nullptr),
                   "modconv_tc: hi/lo outputs must come in pairs");
    MAUA_CHECK_ARG(!ep.noise || ep.noise_weight, "modconv_tc: noise without noise_weight");
  }
  if (batch == 0) return MAUA_OK;

  // halo variant (modconv_tc2.cu) wherever a 16x8 tile fits; v1 keeps the tiny (batch-folded) layers.
  // MAUA_TC_V1=1 forces v1 everywhere (A/B measurements).
  static const bool force_v1 = [] { const char* e = getenv("MAUA_TC_V1"); return e && e[0] == '1'; }();
  if (!force_v1) {
    const int rc2 = modconv_tc2_launch(x_hi, x_lo, w_hi, w_lo, ep, batch, cin, cout, h, w, up, n_products, as_stream(stream));
    if (rc2 != MAUA_E_UNSUPPORTED) return rc2;
  }

  MAUA_CHECK_ARG(!ep.rgb_out, "modconv_tc: fused ToRGB is not available for this shape (needs Cout <= 128, H,W >= 64)");
  MAUA_CHECK_ARG(n_products != 2 && ep.out_fmt == 0,
                 "modconv_tc: the fp16 activation format is only implemented by the halo kernel (GEMM grid >= 64 x 32)");
  Params p;
  p.B = batch; p.H = h; p.W = w; p.Cin = cin; p.Cout = cout;
  p.GH = up ? h + 1 : h;
  p.GW = up ? w + 1 : w;
  // M tile: TB x TH x TW <= 128 pixels (one TMA box; factors need not be powers of two).  Whole images are folded
  // over the batch while they fit; otherwise the (TW, TH) with the fewest tiles wins -- e.g. 17x7 / 11x11 cover the
  // (H+1)^2 grids of the 16^2 / 32^2 up layers with 3 / 9 tiles per image instead of 6 / 15 for 16x8.
  if (p.GW * p.GH <= 128) {
    p.TW = p.GW; p.TH = p.GH;
    p.TB = 128 / (p.GW * p.GH);
    if (p.TB > batch) p.TB = batch;
    const int nb = ceil_div(batch, p.TB);  // same tile count with a more even split (8 images: 5+3 -> 4+4)
    p.TB = ceil_div(batch, nb);
  } else {
    p.TB = 1;
    long long best = -1;
    for (int tw = 1; tw <= 128 && tw <= p.GW; ++tw) {
      int th = 128 / tw;
      if (th > p.GH) th = p.GH;
      const long long n = (long long)ceil_div(p.GW, tw) * ceil_div(p.GH, th);
      if (best < 0 || n < best || (n == best && tw * th >= p.TW * p.TH)) { best = n; p.TW = tw; p.TH = th; }
    }
  }
  p.rows = p.TW * p.TH * p.TB;
  p.tiles_x = ceil_div(p.GW, p.TW);
  p.tiles_y = ceil_div(p.GH, p.TH);
  const int tiles_b = ceil_div(batch, p.TB);
  const long long m_tiles = (long long)p.tiles_x * p.tiles_y * tiles_b;
  const int kc = (cin % 64 == 0) ? 64 : 32;
  p.n_kchunks = cin / kc;
  p.nprod = n_products;
  const int nphase = up ? 4 : 1;
  // (BN, S): modelled cycles = waves * (per-CTA MMA time + fixed prologue/epilogue).  A wide N keeps the MMA efficient
  // (every 128 x N x 16 MMA fetches 4 KB of A whatever N is); when that leaves SMs idle the K loop is split over S
  // CTAs per tile (deterministic reduction through ep.workspace) instead of shrinking N.
  const int n_sm = device_sm_count();
  static const int force_s = [] { const char* e = getenv("MAUA_TC_SPLITK"); return e ? atoi(e) : 0; }();
  int bn = 16, best_s = 1;
  double best_cost = 1e30;
  for (int c = 256; c >= 16; c >>= 1) {
    if (cout % c != 0 || c * nphase > 512) continue;
    for (int sk = 1; sk <= p.n_kchunks && sk <= 16; sk <<= 1) {
      if (p.n_kchunks % sk != 0) continue;
      if (force_s && sk != force_s && !(force_s > p.n_kchunks && sk == 1)) continue;
      const long long ctas = m_tiles * (cout / c) * sk;
      if (sk > 1) {
        const long long need = 4096 + m_tiles * (cout / c) * sk * (long long)(nphase * c) * 128 * 4;
        if (!ep.workspace || need > ep.workspace_bytes || m_tiles * (cout / c) > 1024) continue;
      }
      const double fetch = (4096.0 + 32.0 * c) / 115.0, mma = (c / 2.0 > fetch ? c / 2.0 : fetch);
      const double per_cta = 9.0 * (p.n_kchunks / sk) * (kc / 16) * (n_products > 1 ? 3 : 1) * mma + 6000.0 +
                             (sk > 1 ? 4000.0 + 1000.0 * ((sk + 3) / 4) * (nphase * c / 16) : 0.0);  // reduce: L2 latency bound
      const double cost = (double)ceil_div(ctas, (long long)n_sm) * per_cta;
      if (cost < best_cost) { best_cost = cost; bn = c; best_s = sk; }
    }
  }
  p.BN = bn;
  p.S = best_s;
  p.n_tiles = cout / bn;
  int cols = 32;
  while (cols < bn * nphase) cols <<= 1;
  p.tmem_cols = (uint32_t)cols;
  const int a_stage = 2 * 128 * kc * 2, b_stage = 2 * bn * kc * 2;
  // short-K layers (high resolution, few channels) are latency-bound per CTA: keep the ring small enough that
  // several CTAs co-reside on an SM and overlap each other's prologue / epilogue with MMA work
  const int ctas_per_sm = (p.n_kchunks <= 2) ? 4 : (p.n_kchunks <= 4 ? 2 : 1);
  int stages = ((216 * 1024) / ctas_per_sm - 2048) / (a_stage + b_stage);
  if (stages > 6) stages = 6;
  if (stages < 2) stages = 2;
  p.SA = stages;
  p.SB = stages;
  static const bool debug = [] { const char* e = getenv("MAUA_TC_DEBUG"); return e && e[0] == '1'; }();
  if (debug)
    fprintf(stderr, "[modconv_tc v1] %s B%d %d->%d @%dx%d: tile %dx%dx%d (%d rows) m_tiles=%lld BN=%d S=%d stages=%d grid=%lld\n",
            up ? "up" : "same", batch, cin, cout, h, w, p.TB, p.TH, p.TW, p.rows, m_tiles, bn, p.S, stages,
            m_tiles * p.n_tiles * p.S);
  set_conv_config("v1 up=%d KC=%d tile=%dx%dx%d BN=%d S=%d stages=%d grid=%lld", up, kc, p.TB, p.TH, p.TW, bn, p.S, stages,
                  m_tiles * p.n_tiles * p.S);
  const size_t smem = (size_t)stages * (a_stage + b_stage) + 8 * (2 * p.SA + 2 * p.SB + 2) + 1024;
  MAUA_CHECK_ARG(smem <= 227 * 1024, "modconv_tc: shared memory budget exceeded");
  MAUA_CHECK_ARG(m_tiles * p.n_tiles * p.S < (1LL << 31), "modconv_tc: grid too large");

  // tensor maps: activations [B,H,W,Cin] (dims innermost first), weights [9][Cout][Cin]
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  const cuuint64_t adims[4] = {(cuuint64_t)cin, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)batch};
  const cuuint64_t astr[3] = {(cuuint64_t)cin * 2, (cuuint64_t)w * cin * 2, (cuuint64_t)h * w * cin * 2};
  const cuuint32_t abox[4] = {(cuuint32_t)kc, (cuuint32_t)p.TW, (cuuint32_t)p.TH, (cuuint32_t)p.TB};
  const cuuint64_t bdims[3] = {(cuuint64_t)cin, (cuuint64_t)cout, 9};
  const cuuint64_t bstr[2] = {(cuuint64_t)cin * 2, (cuuint64_t)cout * cin * 2};
  const cuuint32_t bbox[3] = {(cuuint32_t)kc, (cuuint32_t)bn, 1};
  int rc;
  if ((rc = encode_bf16(&ta_hi, x_hi, 4, adims, astr, abox, kc * 2)) != MAUA_OK) return rc;
  if ((rc = encode_bf16(&tb_hi, w_hi, 3, bdims, bstr, bbox, kc * 2)) != MAUA_OK) return rc;
  if (n_products > 1) {
    if ((rc = encode_bf16(&ta_lo, x_lo, 4, adims, astr, abox, kc * 2)) != MAUA_OK) return rc;
    if ((rc = encode_bf16(&tb_lo, w_lo, 3, bdims, bstr, bbox, kc * 2)) != MAUA_OK) return rc;
  } else {
    ta_lo = ta_hi;
    tb_lo = tb_hi;
  }

  cudaStream_t st = as_stream(stream);
  const unsigned grid = (unsigned)(m_tiles * p.n_tiles * p.S);
#define MAUA_TC_LAUNCH(KCV, UPV)                                                                                   \
  do {                                                                                                             \
    MAUA_CHECK_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(modconv_tc_kernel<KCV, UPV>), smem));             \
    MAUA_CHECK_CUDA(launch_chain(modconv_tc_kernel<KCV, UPV>, dim3(grid), dim3(192), smem, st, 1, ta_hi, ta_lo, tb_hi,  \
                                 tb_lo, p, ep));                                                                   \
  } while (0)
  if (kc == 64) {
    if (up) MAUA_TC_LAUNCH(64, true); else MAUA_TC_LAUNCH(64, false);
  } else {
    if (up) MAUA_TC_LAUNCH(32, true); else MAUA_TC_LAUNCH(32, false);
  }
#undef MAUA_TC_LAUNCH
  MAUA_CHECK_LAUNCH("modconv_tc");
  return MAUA_OK;
}
