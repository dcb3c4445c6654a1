// ModulatedConv2d 3x3 on tcgen05 — "halo" variant (v2): ONE activation load per K-chunk serves all 9 taps.
//
// v1 (modconv_tc.cu) issues one shifted TMA box per tap, so the A operand crosses L2->SMEM 9 times (4 times for the
// transposed conv) and the ncu profile of round 1 shows the >=128^2 layers bound by L2->SM traffic, not by the tensor
// pipe.  Here a CTA owns R vertically stacked tiles of 16 rows x 8 columns (M = 128 each).  Per K-chunk it loads the
// (16R+2) x 10 pixel halo ONCE (4-D TMA box, zero padding = out-of-bounds fill) and every tap's A operand is just a
// different START ADDRESS into that halo:   start = halo + ((r*16 + dy) * HW + dx) * row_bytes,   SBO = HW * row_bytes
// (8 consecutive M rows = 8 consecutive pixels of one image row = contiguous shared-memory rows; the next 8-row group
// is the next image row, one halo pitch further).  The 64B/128B swizzle is a function of absolute shared-memory address
// bits, so TMA writes and UMMA reads stay consistent for any row offset and any group pitch (verified on B200 with
// tools/exp_shifted_desc.cu).  The weight tile of a tap is loaded once and used by all R accumulators.
// A traffic per output tile: (16R+2)*10 / (128R) = 1.3-1.4x instead of 9x; B traffic: 1/R.
//
// Everything else (split-bf16 3-product MMA, TMEM accumulators, epilogue fusion, sub-pixel phases for the transposed
// conv) is as in v1.  Warp roles: warp 0 TMA producer, warp 1 TMEM owner + MMA issuer, warps 2-5 epilogue.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include <cmath>

#include "common.cuh"
#include "sm100_ptx.cuh"
#include "tmap.cuh"

namespace maua {
namespace tc2 {

constexpr int TH = 16, TW = 8;  // one M tile = 16 rows x 8 columns
// Epilogue warps = 4 * EPI_GROUPS (each TMEM lane quadrant is served by EPI_GROUPS warps).
//   EPI_GROUPS = 3 (default): 2 role warps + 12 epilogue warps = 448 threads, 128 registers each.
//   EPI_GROUPS = 4 (-DMAUA_EPI_GROUPS=4): the role warps get a warpgroup of their own (warps 0-3) and hand registers to the
//   16 epilogue warps with setmaxnreg (a plain 640-thread launch caps every thread at 96 registers).
#ifndef MAUA_EPI_GROUPS
#define MAUA_EPI_GROUPS 3
#endif
constexpr int EPI_GROUPS = MAUA_EPI_GROUPS;
#ifndef MAUA_TC_PAIR_DEFAULT
#define MAUA_TC_PAIR_DEFAULT 1   // CTA pairs: 1 = on for the BN >= 128 split-bf16 layers (MAUA_TC_PAIR=0 turns them off, =2: BN >= 32)
#endif
#ifndef MAUA_ROLE_WARPS
#define MAUA_ROLE_WARPS (MAUA_EPI_GROUPS >= 4 ? 4 : 2)
#endif
constexpr int ROLE_WARPS = MAUA_ROLE_WARPS;   // warps before the first epilogue warp (4: a warpgroup of their own -> setmaxnreg)
constexpr int THREADS = 32 * ROLE_WARPS + 128 * EPI_GROUPS;

struct Params {
  int B, H, W, Cin, Cout;
  int GH, GW;            // GEMM pixel grid: (H, W) same-res, (H+1, W+1) transposed
  int R;                 // stacked M tiles per CTA
  int HW_, HH_;          // halo width / height in pixels: same-res (10, 16R+2); up (9, 16R+1)
  int tiles_x, tiles_y;  // tiles_y counts groups of R tiles
  int BN, n_tiles;
  int n_kchunks;
  int SA, SB;
  int a_planes;          // activation planes per A stage: 2 = bf16 (hi, lo) pair; 1 = single plane (bf16 fast mode / fp16)
  int b_planes;          // weight planes per B stage: 2 = (hi, lo) pair, 1 = hi only
  int cat;               // 1: [w_hi | w_lo] consumed as ONE MMA of N = 2*BN (accumulator halves added by the epilogue)
  uint32_t a_plane;      // bytes of one bf16 plane of the halo, rounded up to 1024
  uint32_t tmem_cols;
  int AS;                // TMEM accumulator stages (2 = epilogue of item i overlaps the MMAs of item i+1)
  int resident_b;        // 1: all 9*n_kchunks weight tiles fit the B ring: loaded once per CTA, never released
  int n_phase;           // accumulator phases per item: 1 same-res; transposed: 4 / 2 / 1 (4 / n_groups)
  int n_groups;          // transposed only: phase groups walked as separate work items (1, 2 or 4)
  int n_items;           // work items = n_tiles * tiles_x * tiles_y * B, walked with stride gridDim.x
  // exact division of item indices by runtime constants without the ~35-instruction integer-division sequences (every
  // role warp decodes every item: 160 of the ~1000 epilogue instructions per item were these, ncu source view)
  uint32_t fd_m[4], fd_s[4];   // 0: per_group, 1: n_tiles, 2: tiles_x, 3: tiles_y   q = (umulhi(x, m) + x) >> s, x < 2^31
  int per_group;
  // CTA pairs (template PAIR): a work item is one (n tile, phase group) for TWO pixel tiles — cluster rank r takes pixel
  // tile 2*pp + r of pixel-pair pp (n_ptiles odd: the last pair's rank 1 works on an all-out-of-range tile)
  int pair;
  int n_ptiles;          // pixel tiles = tiles_x * tiles_y * B
  int coll;              // transposed conv, fp16 format, one phase group: taps collapsed by input shift (see c_shifts)
  int dbg;               // MAUA_TC_DBG (timing experiments, results invalid): 1 = empty epilogue, 2 = no A loads after the
                         // first, 4 = epilogue stops after the TMEM loads, 8 = epilogue without TMEM loads
};

// Tap lists.  same-res: tap (ky,kx) reads x[y+ky-1, x+kx-1] -> halo (ky, kx), one accumulator phase.
// transposed: u[2y'+py, 2x'+px] += W[ky,kx] * x[y'+dy, x'+dx], py = ky&1, dy = (ky==2 ? -1 : 0) -> halo row dy+1;
// the four sub-pixel phases (py,px) accumulate side by side in TMEM and share the one halo load.
struct Tap { int8_t hy, hx, tap, phase; };  // phase = accumulator index LOCAL to the work item
struct TapList { int32_t n; Tap t[9]; };
// [0] same-res; transposed: [1] all four phases in one item; [2..3] two phase groups {0,3} / {1,2} (5 + 4 taps: the
// most even split of the 4+2+2+1 taps); [4..7] one phase per item.  Phase splitting trades extra halo loads (cheap) for
// fewer accumulator columns per item, i.e. room for two accumulator stages (epilogue overlap) and/or a wider N tile.
// Items are ordered group-slowest, so every persistent CTA walks the same mix of heavy and light groups.
__constant__ TapList c_taps[8] = {
    {9, {{0, 0, 0, 0}, {0, 1, 1, 0}, {0, 2, 2, 0}, {1, 0, 3, 0}, {1, 1, 4, 0}, {1, 2, 5, 0}, {2, 0, 6, 0}, {2, 1, 7, 0}, {2, 2, 8, 0}}},
    {9, {{1, 1, 0, 0}, {1, 1, 1, 1}, {1, 0, 2, 0}, {1, 1, 3, 2}, {1, 1, 4, 3}, {1, 0, 5, 2}, {0, 1, 6, 0}, {0, 1, 7, 1}, {0, 0, 8, 0}}},
    {5, {{1, 1, 0, 0}, {1, 0, 2, 0}, {0, 1, 6, 0}, {0, 0, 8, 0}, {1, 1, 4, 1}}},                // phases 0,3
    {4, {{1, 1, 1, 0}, {0, 1, 7, 0}, {1, 1, 3, 1}, {1, 0, 5, 1}}},                              // phases 1,2
    {4, {{1, 1, 0, 0}, {1, 0, 2, 0}, {0, 1, 6, 0}, {0, 0, 8, 0}}},                              // phase 0
    {2, {{1, 1, 1, 0}, {0, 1, 7, 0}}},                                                          // phase 1
    {2, {{1, 1, 3, 0}, {1, 0, 5, 0}}},                                                          // phase 2
    {1, {{1, 1, 4, 0}}}};                                                                       // phase 3

// "Collapsed" transposed conv (fp16 format, all four phases in one item, BN <= 64).  The nine taps read only FOUR distinct
// input shifts (hy, hx): (1,1) feeds all four sub-pixel phases, (1,0) phases {2,0}, (0,1) phases {0,1}, (0,0) phase 0.
// With the phase accumulators of a tile laid out side by side in TMEM in the order [ph2 | ph0 | ph1 | ph3], the taps of
// one shift are ONE MMA of N = n*BN whose B operand is their weight tiles stacked in shared memory (n TMA boxes into one
// stage): 4 MMAs of N = 4BN / 2BN / 2BN / BN per K-step and weight plane instead of 9 of N = BN.  Small-N MMAs are bound by
// the 4 KB A-operand fetch (cycles ~ 32 + N/4), so 64->32 @512: 9*40 = 360 -> 64+48+48+40 = 200 cycles per K-step and plane.
struct Shift { int8_t hy, hx, slot0, n; int8_t tap[4]; };
__constant__ Shift c_shifts[4] = {{1, 1, 0, 4, {3, 0, 1, 4}},
                                  {1, 0, 0, 2, {5, 2, 0, 0}},
                                  {0, 1, 1, 2, {6, 7, 0, 0}},
                                  {0, 0, 1, 1, {8, 0, 0, 0}}};
// accumulator slot -> sub-pixel phase (py*2 + px), one nibble per slot: slots hold [ph2, ph0, ph1, ph3]
constexpr uint32_t COLL_SLOT_PHASE = 0x3102u;

__device__ __forceinline__ float lrelu_s(float v, float slope, float scale) { return (v > 0.f ? v : v * slope) * scale; }

// MODE (bf16 operands): 0 = one product (hi*hi), 1 = three products as three N = BN MMAs, 2 = "concat": hi*[hi|lo] as
//   one N = 2*BN MMA plus lo*hi.
// MODE (fp16 operands, "f16" activation format: ONE fp16 activation plane, weights as an fp16 (hi, lo) pair, so the
//   weight operand is exact to 2^-22 and the only rounding is the 11-bit activation): 3 = a*w_hi + a*w_lo as two
//   N = BN MMAs into the same accumulator, 4 = a*[w_hi|w_lo] as ONE N = 2*BN MMA (BN <= 64: small-N MMAs are bound by the
//   4 KB A-operand fetch, so halving the A planes AND the MMA count nearly halves the tensor-pipe time of those layers).
// A template parameter (not p.cat / p.a_planes) so that the issue loop is branch-free.
// PAIR: two CTAs of a cluster (one TPC) run every MMA together (cta_group::2, M = 256 = the two CTAs' pixel tiles, the
//   weight tile split in halves between them).  The wide layers (N = 256 per MMA) are bound by shared-memory bandwidth
//   with single-CTA MMAs — 4 KB of A + 8 KB of B per 128-cycle MMA plus the TMA refill of the B ring — and a pair
//   halves the B bytes each SM reads and receives.  Only the leader (rank 0) issues MMAs; both load, both drain.
template <int KC, bool UP, int MODE, bool PAIR>
__global__ void __launch_bounds__(THREADS, 1)
modconv_tc2_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
                   const __grid_constant__ CUtensorMap tm_b_hi, const __grid_constant__ CUtensorMap tm_b_lo,
                   const Params p, const MauaConvEpilogue ep) {
  using namespace ptx;
  constexpr uint32_t ROW = KC * 2;
  static_assert(KC == 32, "the issue loop below is written for two K=16 steps per chunk");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_stage = (uint32_t)p.a_planes * p.a_plane;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int bn_cta = PAIR ? p.BN >> 1 : p.BN;    // weight rows this CTA holds per tap / K chunk
  const bool coll = UP && MODE == 3 && !PAIR && p.coll != 0;
  // one weight plane of a B stage: BN rows (half of them per CTA of a pair); collapsed mode stacks up to 4 tap tiles
  const uint32_t b_half = (coll ? 4u * (uint32_t)p.BN : (uint32_t)bn_cta) * ROW, b_stage = 2 * b_half;
  const uint32_t a_base = smem0;
  const uint32_t b_base = a_base + (uint32_t)p.SA * a_stage;
  const uint32_t bar_base = b_base + (uint32_t)p.SB * b_stage;
  const uint32_t a_full = bar_base, a_empty = a_full + 8 * p.SA;
  const uint32_t b_full = a_empty + 8 * p.SA, b_empty = b_full + 8 * p.SB;
  const uint32_t acc_full = b_empty + 8 * p.SB, acc_empty = acc_full + 8 * p.AS;
  const uint32_t tmem_slot = acc_empty + 8 * p.AS;
  // per-sample epilogue vectors [d | bias | s_next | rgb w0 w1 w2], staged once per sample by the epilogue warps
  float* const vec = reinterpret_cast<float*>(smem_raw + (((tmem_slot + 16u + 15u) & ~15u) - smem_u32(smem_raw)));
  // warp index through a shuffle: the compiler then KNOWS it is warp-uniform, so the role branches below are uniform
  // branches and loop state / descriptors of the single-role loops live in uniform registers (no R2UR per MMA)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_a_hi);
    prefetch_tmap(&tm_b_hi);
    if (p.a_planes > 1) prefetch_tmap(&tm_a_lo);
    if (p.b_planes > 1) prefetch_tmap(&tm_b_lo);
    for (int i = 0; i < p.SA; ++i) { mbar_init(a_full + 8 * i, 1); mbar_init(a_empty + 8 * i, 1); }
    for (int i = 0; i < p.SB; ++i) { mbar_init(b_full + 8 * i, 1); mbar_init(b_empty + 8 * i, 1); }
    // (pairs: the epilogue warps of BOTH CTAs release an accumulator stage on the leader's barrier)
    for (int i = 0; i < p.AS; ++i) {
      mbar_init(acc_full + 8 * i, 1);
      mbar_init(acc_empty + 8 * i, (PAIR ? 2 : 1) * 4 * EPI_GROUPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (PAIR) { tmem_alloc_2cta(tmem_slot, p.tmem_cols); tmem_relinquish_2cta(); }
    else { tmem_alloc(tmem_slot, p.tmem_cols); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // the peer's barriers exist before any remote arrive / TMA completion targets them
  tc_fence_after();
  // programmatic dependent launch: everything above touched no global memory
  pdl_launch_dependents();
  pdl_wait();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  // work item -> (phase group, pixel tile, n tile); n fastest (concurrently running CTAs share the halo through L2),
  // group slowest (load balance: the groups have 5/4 or 4/2/2/1 taps)
  auto fdiv = [&](int x, int which) -> int {
    return (int)((__umulhi((uint32_t)x, p.fd_m[which]) + (uint32_t)x) >> p.fd_s[which]);
  };
  auto decode = [&](int item, int& n0, int& grp, int& x0, int& y0, int& b) {
    grp = 0;
    if (UP && p.n_groups > 1) {
      grp = fdiv(item, 0);
      item -= grp * p.per_group;
    }
    int rest = item, n_tile = 0;
    if (p.n_tiles > 1) {
      rest = fdiv(item, 1);
      n_tile = item - rest * p.n_tiles;
    }
    n0 = n_tile * p.BN;
    if (PAIR) {
      rest = 2 * rest + (int)rank;           // pixel-pair -> this CTA's pixel tile
      if (rest >= p.n_ptiles) {              // odd tile count: nothing to do for the last pair's rank 1 — a tile below the
        x0 = 0;                              // image (TMA zero-fills the halo, every epilogue store is bounds-checked)
        y0 = p.tiles_y * TH * p.R;
        b = p.B - 1;
        return;
      }
    }
    const int q1 = fdiv(rest, 2);            // rest / tiles_x
    x0 = (rest - q1 * p.tiles_x) * TW;
    b = fdiv(q1, 3);                         // q1 / tiles_y
    y0 = (q1 - b * p.tiles_y) * TH * p.R;
  };
  // persistent walk: CTA (or CTA pair) i takes items i, i + step, ...
  const int item0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int istep = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto tap_list = [&](int grp) -> const TapList& {
    return c_taps[!UP ? 0 : (p.n_groups == 1 ? 1 : (p.n_groups == 2 ? 2 + grp : 4 + grp))];
  };
  const uint32_t blk_cols = (uint32_t)(p.cat ? 2 * p.BN : p.BN);  // TMEM columns of one accumulator
  const uint32_t acc_cols = (uint32_t)(p.n_phase * p.R) * blk_cols;  // ... of one accumulator stage
  const uint32_t halo_bytes = (uint32_t)(p.HW_ * p.HH_) * ROW;    // bytes one TMA box writes per plane

  if (ROLE_WARPS >= 4) {  // register hand-over between warpgroups (all warps of a warpgroup execute the same instruction)
    if (warp < ROLE_WARPS) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
  }
  if (warp == 0 && lane == 0) {
    // ================================ TMA producer ================================
    int ia = 0, ib = 0;
    uint32_t pa = 0, pb = 0;
    // pairs: every load of either CTA is counted on the LEADER's full barrier (it alone waits for operands); the leader's
    // expect_tx covers both CTAs' boxes, the empty barriers stay per CTA (signalled by the multicast commit)
    const uint32_t a_full_tx = PAIR ? mapa_u32(a_full, 0) : a_full, b_full_tx = PAIR ? mapa_u32(b_full, 0) : b_full;
    const uint32_t ncta = PAIR ? 2u : 1u;
    const int brow = PAIR ? (int)rank * bn_cta : 0;   // this CTA's half of the weight tile
    for (int item = item0; item < p.n_items; item += istep) {
      int n0, grp, x0, y0, b;
      decode(item, n0, grp, x0, y0, b);
      const TapList& tl = tap_list(grp);
      for (int kc = 0; kc < p.n_kchunks; ++kc) {
        const int c0 = kc * KC;
        mbar_wait(a_empty + 8 * ia, pa ^ 1);
        if (!PAIR && (p.dbg & 2) && item != item0) {
          mbar_arrive(a_full + 8 * ia);
        } else {
        if (rank == 0) mbar_expect_tx(a_full + 8 * ia, ncta * halo_bytes * (uint32_t)p.a_planes);
        const uint32_t dst = a_base + ia * a_stage;
        if (PAIR) {
          tma_load_4d_2cta(dst, &tm_a_hi, a_full_tx + 8 * ia, c0, x0 - 1, y0 - 1, b);
          if (p.a_planes > 1) tma_load_4d_2cta(dst + p.a_plane, &tm_a_lo, a_full_tx + 8 * ia, c0, x0 - 1, y0 - 1, b);
        } else {
          tma_load_4d(dst, &tm_a_hi, a_full + 8 * ia, c0, x0 - 1, y0 - 1, b);
          if (p.a_planes > 1) tma_load_4d(dst + p.a_plane, &tm_a_lo, a_full + 8 * ia, c0, x0 - 1, y0 - 1, b);
        }
        }
        if (++ia == p.SA) { ia = 0; pa ^= 1; }
        if (coll) {   // one B stage per input shift: the weight tiles of its taps stacked, hi plane then lo plane
          const uint32_t tile_bytes = (uint32_t)p.BN * ROW;
#pragma unroll 1
          for (int sf = 0; sf < 4; ++sf) {
            const Shift sh = c_shifts[sf];
            mbar_wait(b_empty + 8 * ib, pb ^ 1);
            mbar_expect_tx(b_full + 8 * ib, 2u * (uint32_t)sh.n * tile_bytes);
            const uint32_t dstb = b_base + ib * b_stage;
            for (int j = 0; j < sh.n; ++j) {
              tma_load_3d(dstb + j * tile_bytes, &tm_b_hi, b_full + 8 * ib, c0, n0, sh.tap[j]);
              tma_load_3d(dstb + b_half + j * tile_bytes, &tm_b_lo, b_full + 8 * ib, c0, n0, sh.tap[j]);
            }
            if (++ib == p.SB) { ib = 0; pb ^= 1; }
          }
          continue;
        }
#pragma unroll 1
        for (int t = 0; t < tl.n; ++t) {
          if (p.resident_b && item != item0) break;  // weights already resident in shared memory
          mbar_wait(b_empty + 8 * ib, pb ^ 1);
          if (rank == 0) mbar_expect_tx(b_full + 8 * ib, ncta * (p.b_planes > 1 ? b_stage : b_half));
          const uint32_t dstb = b_base + ib * b_stage;
          if (PAIR) {
            tma_load_3d_2cta(dstb, &tm_b_hi, b_full_tx + 8 * ib, c0, n0 + brow, tl.t[t].tap);
            if (p.b_planes > 1) tma_load_3d_2cta(dstb + b_half, &tm_b_lo, b_full_tx + 8 * ib, c0, n0 + brow, tl.t[t].tap);
          } else {
            tma_load_3d(dstb, &tm_b_hi, b_full + 8 * ib, c0, n0, tl.t[t].tap);
            if (p.b_planes > 1) tma_load_3d(dstb + b_half, &tm_b_lo, b_full + 8 * ib, c0, n0, tl.t[t].tap);
          }
          if (++ib == p.SB) { ib = 0; pb ^= 1; }
        }
      }
    }
  } else if (warp == 1 && (!PAIR || rank == 0)) {
    // ================================ MMA issuer ================================
    // the whole warp walks the loop (uniform control flow); one elected lane issues the tcgen05 instructions
    const bool leader = elect_one_sync();
    // The issue loop must stay far below the ~50 cycles a 128x32x16 MMA occupies the tensor pipe: all descriptor
    // fields are folded into per-stage base values up front; per MMA only 64-bit adds of small constants remain.
    constexpr uint32_t MM = PAIR ? 256u : 128u;
    const uint32_t idesc = MODE >= 3 ? make_idesc_f16(MM, (uint32_t)p.BN) : make_idesc_bf16(MM, (uint32_t)p.BN);
    const uint32_t idesc2 = MODE >= 3 ? make_idesc_f16(MM, (uint32_t)(2 * p.BN)) : make_idesc_bf16(MM, (uint32_t)(2 * p.BN));
    auto mma = [](uint32_t acc, uint64_t da, uint64_t db, uint32_t id, uint32_t accumulate) {
      if (PAIR) umma_bf16_2cta(acc, da, db, id, accumulate);
      else umma_bf16(acc, da, db, id, accumulate);
    };
    auto commit = [](uint32_t bar) {
      if (PAIR) umma_commit_2cta(bar);
      else umma_commit(bar);
    };
    const uint64_t sbo_field = (uint64_t)((((uint32_t)p.HW_ * ROW) >> 4) & 0x3FFF) << 32;
    const uint64_t da0 = (make_kmajor_desc(a_base, ROW) & ~(0x3FFFull << 32)) | sbo_field;  // A stage 0, hi plane
    const uint64_t db0 = make_kmajor_desc(b_base, ROW);                                      // B stage 0, hi plane
    const uint64_t a_stage16 = a_stage >> 4, a_plane16 = p.a_plane >> 4;
    const uint64_t b_stage16 = b_stage >> 4, b_half16 = b_half >> 4;
    const uint32_t rstep16 = (uint32_t)(TH * p.HW_) * ROW >> 4;  // next stacked tile: 16 halo rows further
    constexpr int mode = MODE;
    int ia = 0, ib = 0, as = 0;
    uint32_t pa = 0, pb = 0, pacc = 0;
    for (int item = item0; item < p.n_items; item += istep) {
      int n0, grp, x0, y0, b;
      decode(item, n0, grp, x0, y0, b);
      const TapList& tl = tap_list(grp);
      mbar_wait(acc_empty + 8 * as, pacc ^ 1);  // epilogue has drained this accumulator stage
      tc_fence_after();
      const uint32_t acc_stage = tmem_base + (uint32_t)as * acc_cols;
      uint32_t started = 0;  // bit ph: the accumulators of phase ph have received their first MMA
      // R is a compile-time constant inside the tap loop (generic lambda + switch): the R x (2..6) MMAs of a tap are
      // fully unrolled, so their descriptors sit in distinct uniform registers and the UTCHMMAs issue back to back —
      // with a rolled r loop the warp stalled once per 4 MMAs until the tensor core had consumed the operand registers.
      auto run_item = [&](auto rc) {
        constexpr int RR = decltype(rc)::value;
        for (int kc = 0; kc < p.n_kchunks; ++kc) {
          mbar_wait(a_full + 8 * ia, pa);
          const uint64_t da_stage = da0 + (uint64_t)ia * a_stage16;
          if (coll) {
#pragma unroll 1
            for (int sf = 0; sf < 4; ++sf) {
              const Shift sh = c_shifts[sf];
              mbar_wait(b_full + 8 * ib, pb);
              tc_fence_after();
              const uint64_t dbh = db0 + (uint64_t)ib * b_stage16, dbl = dbh + b_half16;
              const uint64_t dah0 = da_stage + (uint64_t)((uint32_t)(sh.hy * p.HW_ + sh.hx) * ROW >> 4);
              const uint32_t idn = make_idesc_f16(128u, (uint32_t)(sh.n * p.BN));
              // accumulators of tile r: 4 slots of BN columns side by side; the first shift touches all of them
              const uint32_t acc0 = acc_stage + (uint32_t)(sh.slot0 * p.BN);
              const uint32_t accumulate = (kc > 0 || sf > 0) ? 1u : 0u;
              if (leader) {
#pragma unroll
                for (int r = 0; r < RR; ++r) {
                  const uint64_t dah = dah0 + (uint64_t)r * rstep16;
                  const uint32_t acc = acc0 + (uint32_t)(r * 4 * p.BN);
                  mma(acc, dah, dbh, idn, accumulate);
                  mma(acc, dah + 2, dbh + 2, idn, 1u);
                  mma(acc, dah, dbl, idn, 1u);
                  mma(acc, dah + 2, dbl + 2, idn, 1u);
                }
                commit(b_empty + 8 * ib);
              }
              if (++ib == p.SB) { ib = 0; pb ^= 1; }
            }
            if (leader) commit(a_empty + 8 * ia);
            if (++ia == p.SA) { ia = 0; pa ^= 1; }
            continue;
          }
#pragma unroll 1
          for (int t = 0; t < tl.n; ++t) {
            const Tap tp = tl.t[t];
            if (!p.resident_b || item == item0) {
              mbar_wait(b_full + 8 * ib, pb);
              tc_fence_after();
            }
            const uint64_t dbh = db0 + (uint64_t)ib * b_stage16, dbl = dbh + b_half16;
            const uint64_t dah0 = da_stage + (uint64_t)((uint32_t)(tp.hy * p.HW_ + tp.hx) * ROW >> 4);
            const uint32_t acc0 = acc_stage + (uint32_t)(tp.phase * RR) * blk_cols;
            const uint32_t first = (started >> tp.phase) & 1u;
            if (leader) {
#pragma unroll
              for (int r = 0; r < RR; ++r) {
                const uint64_t dah = dah0 + (uint64_t)r * rstep16, dal = dah + a_plane16;
                const uint32_t acc = acc0 + (uint32_t)r * blk_cols;
                if (mode == 2) {  // [hi*hi | hi*lo] in one N = 2*BN MMA (B stage = hi rows then lo rows), then lo*hi
                  mma(acc, dah, dbh, idesc2, first);
                  mma(acc, dah + 2, dbh + 2, idesc2, 1u);
                  mma(acc, dal, dbh, idesc, 1u);
                  mma(acc, dal + 2, dbh + 2, idesc, 1u);
                } else if (mode == 4) {  // fp16: a * [w_hi | w_lo], one N = 2*BN MMA per K-step
                  mma(acc, dah, dbh, idesc2, first);
                  mma(acc, dah + 2, dbh + 2, idesc2, 1u);
                } else if (mode == 3) {  // fp16: a * w_hi + a * w_lo into the same accumulator
                  mma(acc, dah, dbh, idesc, first);
                  mma(acc, dah + 2, dbh + 2, idesc, 1u);
                  mma(acc, dah, dbl, idesc, 1u);
                  mma(acc, dah + 2, dbl + 2, idesc, 1u);
                } else {
                  mma(acc, dah, dbh, idesc, first);
                  mma(acc, dah + 2, dbh + 2, idesc, 1u);
                  if (mode == 1) {
                    mma(acc, dah, dbl, idesc, 1u);
                    mma(acc, dah + 2, dbl + 2, idesc, 1u);
                    mma(acc, dal, dbh, idesc, 1u);
                    mma(acc, dal + 2, dbh + 2, idesc, 1u);
                  }
                }
              }
            }
            started |= 1u << tp.phase;
            if (!p.resident_b && leader) commit(b_empty + 8 * ib);
            if (++ib == p.SB) { ib = 0; pb ^= 1; }
          }
          if (leader) commit(a_empty + 8 * ia);
          if (++ia == p.SA) { ia = 0; pa ^= 1; }
        }
      };
      if (p.R == 4) run_item(std::integral_constant<int, 4>{});
      else if (p.R == 2) run_item(std::integral_constant<int, 2>{});
      else run_item(std::integral_constant<int, 1>{});
      if (leader) commit(acc_full + 8 * as);
      if (++as == p.AS) { as = 0; pacc ^= 1; }
    }
  } else if (warp >= ROLE_WARPS) {
    // ================================ epilogue ================================
    // 4*EPI_GROUPS warps: warp w serves TMEM lane quadrant (w & 3); the 16-column chunks of all accumulators of an
    // item are dealt round-robin to the EPI_GROUPS warps of a quadrant (a lone warp per scheduler exposes every latency)
    const int quad = warp & 3;
    const int egroup = (warp - ROLE_WARPS) >> 2;
    const int m = quad * 32 + lane;
    const int tx = m & (TW - 1), ty = m >> 3;
    // activation gain folded into the staged vectors (see the staging loop); slope outside [0, 1] keeps the generic form
    const bool fold_gain = ep.activate != 0 && ep.act_scale > 0.f && ep.slope >= 0.f && ep.slope <= 1.f;
    const float gain = fold_gain ? ep.act_scale : 1.f;
    const float nwv = (!UP && ep.activate && ep.noise) ? __ldg(ep.noise_weight) * gain : 0.f;
    int as = 0;
    uint32_t pacc = 0;
    // Hot-loop parameters pinned in registers: ptxas otherwise re-reads them from the constant bank inside the job loop
    // and the LDC latency was 25 % of the epilogue's stall samples (ncu, 64->32 @512 up).
    int R = p.R, BN = p.BN, Cout = p.Cout, n_phase = p.n_phase, GH = p.GH, GW = p.GW;
    asm volatile("" : "+r"(R), "+r"(BN), "+r"(Cout), "+r"(n_phase), "+r"(GH), "+r"(GW));
    const int bn16_log = 31 - __clz(BN >> 4), r_log = 31 - __clz(R);  // BN/16 and R are powers of two
    const int nchunk = BN >> 4, ntile = n_phase * R;
    const bool cat = p.cat != 0, act = ep.activate != 0;
    // fused ToRGB (host guarantees BN == Cout): one warp owns all chunks of a tile so that the three dot products are
    // summed in a fixed order; tile r of the CTA's it-th item goes to epilogue group (it*R + r) % EPI_GROUPS, so that
    // R = 2 / R = 4 items load the three groups evenly over consecutive items.  Otherwise a job is one 16-column chunk
    // of one accumulator, dealt round-robin.  Every warp walks only ITS jobs (no skip iterations).
    const bool fuse_rgb = !UP && ep.rgb_out != nullptr;
    const int OH = UP ? 2 * p.H + 1 : p.H, OW = UP ? 2 * p.W + 1 : p.W;
    // Per-channel epilogue operands (demodulation d[b,:], bias, next-layer style s_next[b,:], ToRGB rows) are the same
    // for every pixel of a sample.  Read through __ldg per 16-column chunk they cost 8-20 L1 round trips whose latency
    // the three warps per scheduler cannot hide (ncu: long-scoreboard stalls on their first use = 45 % of the epilogue
    // samples, plus spills from holding them across the TMEM wait).  They are staged in shared memory whenever the
    // CTA's item sequence moves to another sample (batch is the slowest item coordinate: <= batch times per launch).
    float* const sm_d = vec;
    float* const sm_b = vec + Cout;
    float* const sm_s = vec + 2 * Cout;
    float* const sm_w = vec + 3 * Cout;
    int cur_b = -1;
    int job0 = 0;
    const uint32_t acc_empty_tx = PAIR ? mapa_u32(acc_empty, 0) : acc_empty;   // the leader's barrier
    for (int item = item0; item < p.n_items; item += istep, job0 = (job0 + R) % EPI_GROUPS) {
      int n0, grp, x0, y0, b;
      decode(item, n0, grp, x0, y0, b);
      if (!UP && b != cur_b) {  // uniform over the epilogue warps: they all walk the same item sequence
        asm volatile("bar.sync 1, %0;" ::"n"(128 * EPI_GROUPS) : "memory");  // nobody still reads the old vectors
        const int et = (int)threadIdx.x - 32 * ROLE_WARPS;
        // lrelu(x) * g = max(x * g, slope * x * g) for g > 0, 0 <= slope <= 1: the activation gain g is folded into the
        // staged demodulation / bias vectors (and the noise scalar), which removes one multiply per element and turns the
        // compare + select pairs into single FMNMX instructions (the epilogue, not the tensor pipe, paces the Cout <= 64
        // layers: ncu, 32->32 @1024^2, 0.64 eligible warps per scheduler)
        for (int i = et; i < Cout; i += 128 * EPI_GROUPS) {
          sm_d[i] = (ep.d ? __ldg(ep.d + (long long)b * Cout + i) : 1.f) * gain;
          sm_b[i] = (act && ep.bias) ? __ldg(ep.bias + i) * gain : 0.f;
          sm_s[i] = ep.s_next ? __ldg(ep.s_next + (long long)b * Cout + i) : 1.f;
        }
        if (fuse_rgb)
          for (int i = et; i < 3 * Cout; i += 128 * EPI_GROUPS) sm_w[i] = __ldg(ep.rgb_w + (long long)b * 3 * Cout + i);
        asm volatile("bar.sync 1, %0;" ::"n"(128 * EPI_GROUPS) : "memory");
        cur_b = b;
      }
      const int gx = x0 + tx;
      // Noise of this lane's pixel in each of the item's R tiles, requested BEFORE waiting for the accumulators: the
      // noise map is streamed from HBM/L2 (4 B per pixel, no reuse), and with the load next to its use its ~1 us latency
      // was the largest single stall of the same-resolution epilogue (ncu: 21 % of its samples on the consuming FMUL).
      // (with fused ToRGB a warp owns whole tiles — first_job and first_job + EPI_GROUPS — and only needs their values)
      const int first_job_rgb = (egroup + EPI_GROUPS - job0) % EPI_GROUPS;
      float nzv[4] = {0.f, 0.f, 0.f, 0.f};
      if (!UP && act && ep.noise && gx < GW) {
        // one 64-bit address per item, the R tiles are TH rows apart (a 32-bit constant step)
        const float* np = ep.noise + (long long)b * ep.noise_bstride + (long long)(y0 + ty) * OW + gx;
        const int rstep = TH * OW;
        if (fuse_rgb) {
          // this warp's (at most two) tiles: nzv[0] / nzv[1] = noise of its first / second job (two address
          // computations and loads instead of four predicated ones: ~90 of the ~750 instructions per item and warp)
          const int ra = first_job_rgb, rb = first_job_rgb + EPI_GROUPS;
          if (ra < R && y0 + ra * TH + ty < GH) nzv[0] = __ldg(np + ra * rstep);
          if (rb < R && y0 + rb * TH + ty < GH) nzv[1] = __ldg(np + rb * rstep);
        } else {
#pragma unroll
          for (int r = 0; r < 4; ++r)
            if (r < R && y0 + r * TH + ty < GH) nzv[r] = __ldg(np + r * rstep);
        }
      }
      const long long pix_item = UP ? ((long long)b * OH + 2 * (y0 + ty)) * OW + 2 * gx
                                    : ((long long)b * OH + (y0 + ty)) * OW + gx;
      mbar_wait(acc_full + 8 * as, pacc);
      tc_fence_after();
      const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)as * acc_cols;
      // (the transposed conv only needs d: 4 cached loads per chunk, issued before the TMEM load — measured faster than
      // the shared-memory copy there, whose reads compete with the UMMA operand fetches of its busier tensor pipe)
      const float* dglob = (UP && ep.d) ? ep.d + (long long)b * Cout + n0 : nullptr;
      const float4* dptr = reinterpret_cast<const float4*>(sm_d + n0);
      const float4* bptr = reinterpret_cast<const float4*>(sm_b + n0);
      const float4* sptr = reinterpret_cast<const float4*>(sm_s + n0);
      const int njobs = (p.dbg & 1) ? 0 : (fuse_rgb ? R : ntile * nchunk);
      const int first_job = fuse_rgb ? first_job_rgb : egroup;
#pragma unroll 1
      for (int job = first_job; job < njobs; job += EPI_GROUPS) {
        const int tile = fuse_rgb ? job : (job >> bn16_log);   // accumulator index = ph * R + r (collapsed: r * 4 + slot)
        const int ph = coll ? (int)((COLL_SLOT_PHASE >> (4 * (tile & 3))) & 3u) : tile >> r_log;
        const int r = coll ? tile >> 2 : tile & (R - 1);
        const int c_first = fuse_rgb ? 0 : ((job & (nchunk - 1)) << 4);
        const int c_end = fuse_rgb ? BN : c_first + 16;
        const int gy = y0 + r * TH + ty;
        const bool in_grid = (gy < GH) && (gx < GW);
        // global sub-pixel phase (py, px) = (gph >> 1, gph & 1); two groups hold phases {0,3} and {1,2}
        const int gph = !UP ? 0 : (p.n_groups == 2 ? (grp == 0 ? 3 * ph : 1 + ph) : grp * n_phase + ph);
        const int oy = UP ? 2 * gy + (gph >> 1) : gy, ox = UP ? 2 * gx + (gph & 1) : gx;
        const bool valid = in_grid && oy < OH && ox < OW;
        // output pixel index = per-item base (one 64-bit product per item) + a 32-bit offset per job
        const long long pix = valid ? pix_item + (UP ? (2 * r * TH + (gph >> 1)) * OW + (gph & 1) : r * TH * OW) : 0;
        const float nz = UP ? 0.f
                            : nwv * (fuse_rgb ? (job == first_job ? nzv[0] : nzv[1])
                                              : (r == 0 ? nzv[0] : (r == 1 ? nzv[1] : (r == 2 ? nzv[2] : nzv[3]))));
        const uint32_t acc_col = (uint32_t)tile * blk_cols;
        float2 rgb0 = make_float2(0.f, 0.f), rgb1 = rgb0, rgb2 = rgb0;
#pragma unroll 1
        for (int c = c_first; c < c_end; c += 16) {   // a single iteration unless ToRGB is fused
          // transposed conv: demodulation is normally left to maua_blur_act_nhwc (it commutes with the per-channel FIR),
          // so the raw phases go straight from TMEM to HBM; ep.d != NULL keeps the multiply here (stand-alone use)
          float4 dup[4];
          if (UP && dglob) {
#pragma unroll
            for (int q = 0; q < 4; ++q) dup[q] = __ldg(reinterpret_cast<const float4*>(dglob + c) + q);
          }
          uint32_t rr[16];
          if (p.dbg & 8) {
#pragma unroll
            for (int i = 0; i < 16; ++i) rr[i] = (uint32_t)(c + i + lane);
          } else {
            tmem_ld_x16(lane_addr + acc_col + (uint32_t)c, rr);
          }
          // per-channel operands of this chunk, requested while the TMEM loads are in flight
          float4 dv[4], bv[4];
          if (!UP) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              dv[q] = dptr[(c >> 2) + q];
              bv[q] = bptr[(c >> 2) + q];
            }
          }
          float2 v[8];
          if (cat && !(p.dbg & 8)) {
            uint32_t r2[16];
            tmem_ld_x16(lane_addr + acc_col + (uint32_t)(BN + c), r2);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i)
              v[i] = fadd2(make_float2(__uint_as_float(rr[2 * i]), __uint_as_float(rr[2 * i + 1])),
                           make_float2(__uint_as_float(r2[2 * i]), __uint_as_float(r2[2 * i + 1])));
          } else {
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = make_float2(__uint_as_float(rr[2 * i]), __uint_as_float(rr[2 * i + 1]));
          }
          if (!valid || (p.dbg & 4)) continue;
          // Packed fp32 (fma/add/mul.rn.f32x2: two IEEE-rounded results per issue slot): the epilogue's instruction
          // stream, not the tensor pipe, paced the Cout <= 128 layers (ncu: 263 SASS instructions per 16-column chunk,
          // 160 of them scalar FADD/FMUL/FFMA) — the pairs below halve that part.
          if (UP) {
            if (dglob) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                v[2 * q] = fmul2(v[2 * q], make_float2(dup[q].x, dup[q].y));
                v[2 * q + 1] = fmul2(v[2 * q + 1], make_float2(dup[q].z, dup[q].w));
              }
            }
            // 64 contiguous bytes per lane as TWO 256-bit stores: whole 32-byte sectors per request (four 128-bit stores
            // produced half-sector writes: ncu counted 2x the ideal L2 store sectors on the 64->32 @512 up layer)
            float* dst = ep.out_raw_nhwc + pix * Cout + n0 + c;
            st_global_v8(dst, v[0], v[1], v[2], v[3]);
            st_global_v8(dst + 8, v[4], v[5], v[6], v[7]);
          } else {
            // o = lrelu(acc * d + noise + bias) * sqrt2: acc*d + noise as one FMA (one rounding less than the reference's
            // separate multiply and add; the tensor-core path is tolerance-checked, 1e-3), then + bias
            const float2 nz2 = make_float2(nz, nz);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              v[2 * q] = ffma2(v[2 * q], make_float2(dv[q].x, dv[q].y), nz2);
              v[2 * q + 1] = ffma2(v[2 * q + 1], make_float2(dv[q].z, dv[q].w), nz2);
              if (act) {
                v[2 * q] = fadd2(v[2 * q], make_float2(bv[q].x, bv[q].y));
                v[2 * q + 1] = fadd2(v[2 * q + 1], make_float2(bv[q].z, bv[q].w));
              }
            }
            if (act && fold_gain) {   // gain already inside d / bias / noise: lrelu = max(x, slope * x)
              const float2 sl2 = make_float2(ep.slope, ep.slope);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float2 t = fmul2(v[i], sl2);
                v[i] = make_float2(fmaxf(v[i].x, t.x), fmaxf(v[i].y, t.y));
              }
            } else if (act) {
              const float2 sl2 = make_float2(ep.slope, ep.slope), sc2 = make_float2(ep.act_scale, ep.act_scale);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float2 t = fmul2(v[i], sl2);
                v[i] = fmul2(make_float2(v[i].x > 0.f ? v[i].x : t.x, v[i].y > 0.f ? v[i].y : t.y), sc2);
              }
            }
            if (fuse_rgb) {
              // three dot products over the tile's channels; (even, odd) channel partial sums ride in one register
              // pair and are combined when the tile is stored (fixed order: deterministic)
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 w0 = reinterpret_cast<const float4*>(sm_w + c)[q];
                const float4 w1 = reinterpret_cast<const float4*>(sm_w + Cout + c)[q];
                const float4 w2 = reinterpret_cast<const float4*>(sm_w + 2 * Cout + c)[q];
                rgb0 = ffma2(v[2 * q], make_float2(w0.x, w0.y), rgb0);
                rgb0 = ffma2(v[2 * q + 1], make_float2(w0.z, w0.w), rgb0);
                rgb1 = ffma2(v[2 * q], make_float2(w1.x, w1.y), rgb1);
                rgb1 = ffma2(v[2 * q + 1], make_float2(w1.z, w1.w), rgb1);
                rgb2 = ffma2(v[2 * q], make_float2(w2.x, w2.y), rgb2);
                rgb2 = ffma2(v[2 * q + 1], make_float2(w2.z, w2.w), rgb2);
              }
            }
            if (ep.out_f32_nchw) {
              float* dst = ep.out_f32_nchw + (((long long)b * Cout + n0 + c) * OH + oy) * OW + ox;
              const long long plane = (long long)OH * OW;
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                dst[(2 * i) * plane] = v[i].x;
                dst[(2 * i + 1) * plane] = v[i].y;
              }
            }
            if (ep.out_hi && ep.out_fmt == 1) {
              // "f16" activation format for the consumer: ONE fp16 plane of o * s_next (saturating conversion)
              uint32_t h[8];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 sv = sptr[(c >> 2) + q];
                const float2 a01 = fmul2(v[2 * q], make_float2(sv.x, sv.y)), a23 = fmul2(v[2 * q + 1], make_float2(sv.z, sv.w));
                h[2 * q] = pack_f16x2_sat(a01.x, a01.y);
                h[2 * q + 1] = pack_f16x2_sat(a23.x, a23.y);
              }
              // 16 channels x 2 B = one 32-byte sector per pixel: a single 256-bit store
              st_global_v8_b32(reinterpret_cast<__half*>(ep.out_hi) + pix * Cout + n0 + c, h[0], h[1], h[2], h[3], h[4], h[5],
                               h[6], h[7]);
            } else if (ep.out_hi) {
              uint32_t h[8], l[8];
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float4 sv = sptr[(c >> 2) + q];
                const float2 a01 = fmul2(v[2 * q], make_float2(sv.x, sv.y)), a23 = fmul2(v[2 * q + 1], make_float2(sv.z, sv.w));
                const __nv_bfloat162 h01 = __floats2bfloat162_rn(a01.x, a01.y), h23 = __floats2bfloat162_rn(a23.x, a23.y);
                const float2 f01 = __bfloat1622float2(h01), f23 = __bfloat1622float2(h23);
                const float2 neg1 = make_float2(-1.f, -1.f);
                const float2 r01 = ffma2(f01, neg1, a01), r23 = ffma2(f23, neg1, a23);   // exact: a - hi
                const __nv_bfloat162 l01 = __floats2bfloat162_rn(r01.x, r01.y), l23 = __floats2bfloat162_rn(r23.x, r23.y);
                h[2 * q] = *reinterpret_cast<const uint32_t*>(&h01);
                h[2 * q + 1] = *reinterpret_cast<const uint32_t*>(&h23);
                l[2 * q] = *reinterpret_cast<const uint32_t*>(&l01);
                l[2 * q + 1] = *reinterpret_cast<const uint32_t*>(&l23);
              }
              st_global_v8_b32(reinterpret_cast<__nv_bfloat16*>(ep.out_hi) + pix * Cout + n0 + c, h[0], h[1], h[2], h[3], h[4],
                               h[5], h[6], h[7]);
              st_global_v8_b32(reinterpret_cast<__nv_bfloat16*>(ep.out_lo) + pix * Cout + n0 + c, l[0], l[1], l[2], l[3], l[4],
                               l[5], l[6], l[7]);
            }
          }
        }
        if (fuse_rgb && gy < p.H && gx < p.W) {
          float* o = ep.rgb_out + (((long long)b * 3) * p.H + gy) * p.W + gx;
          const long long plane = (long long)p.H * p.W;
          o[0] = rgb0.x + rgb0.y; o[plane] = rgb1.x + rgb1.y; o[2 * plane] = rgb2.x + rgb2.y;
        }
      }
      // accumulator stage drained: hand it back to the MMA issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(acc_empty_tx + 8 * as);
        else mbar_arrive(acc_empty + 8 * as);
      }
      if (++as == p.AS) { as = 0; pacc ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // neither CTA leaves (or frees TMEM) while the other may still signal / read it
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_2cta(tmem_base, p.tmem_cols);
    else tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

static inline uint32_t align1k(uint32_t v) { return (v + 1023u) & ~1023u; }

}  // namespace tc2

// Returns MAUA_E_UNSUPPORTED when the shape is better served by v1 (tiny images: batch-folded tiles).
int modconv_tc2_launch(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo,
                       const MauaConvEpilogue& ep, int batch, int cin, int cout, int h, int w, int up, int n_products,
                       cudaStream_t st) {
  using namespace tc2;
  const int GH = up ? h + 1 : h, GW = up ? w + 1 : w;
  // MAUA_TC_FORCE="R,BN,cat,groups" (read on every call: tools/tune_tc2.py and the unit tests sweep it) pins the
  // configuration and bypasses the size gate below; an infeasible forced configuration is an error, never a fallback.
  int f_r = 0, f_bn = 0, f_cat = 0, f_groups = 0;
  if (const char* f = getenv("MAUA_TC_FORCE")) {
    if (sscanf(f, "%d,%d,%d,%d", &f_r, &f_bn, &f_cat, &f_groups) != 4) f_r = 0;
  }
  // <= 32^2 (measured): v1's batch-folded tiles are faster — except the same-resolution Cout >= 256 layer at 32^2 once
  // the batch fills the machine with (R=1, BN=256) items (512->512 @32^2, batch 8: 0.088 ms vs 0.103 ms for v1).
  // MAUA_TC2_MIN_TILES overrides the threshold (experiments).
  static const int min_tiles = [] { const char* e = getenv("MAUA_TC2_MIN_TILES"); return e ? atoi(e) : 4; }();
  if (!f_r && (GH < min_tiles * TH || GW < min_tiles * TW)) {
    const bool wide32 = !up && GH >= 2 * TH && GW >= 2 * TW && cout >= 256 && cout % 256 == 0 &&
                        (long long)ceil_div(GW, TW) * ceil_div(GH, TH) * batch * (cout / 256) >= 120;
    if (!wide32) return MAUA_E_UNSUPPORTED;
  }
  Params p;
  p.B = batch; p.H = h; p.W = w; p.Cin = cin; p.Cout = cout; p.GH = GH; p.GW = GW;
  const int nphase = up ? 4 : 1;
  const int kc = 32;  // 64-byte operand rows: the halo of R = 4 stacked tiles still fits next to a deep B ring
  const int n_kchunks = cin / kc;
  const long long tiles_x = ceil_div(GW, TW);
  const long long rows16 = ceil_div(GH, TH);
  // Persistent kernel: one CTA per SM owns the whole shared memory (deep rings) and all 512 TMEM columns; when two
  // accumulator stages fit (2*R*nphase*BN <= 512) the epilogue of item i overlaps the MMAs of item i+1.
  const int tmem_cap = 512;
  const uint32_t budget = 212u * 1024u - 6u * (uint32_t)cout * 4u;  // minus the epilogue's per-sample vectors
  // Configuration (R stacked tiles, BN, concat mode, phase groups).  tools/tune_tc2.py sweeps the whole space on
  // hardware; across config-f 1024^2 (batch 2 and 8) and config-e 512^2 the winner only depends on Cout and on the
  // layer kind -- widest N first (a 128 x N x 16 MMA fetches its 4 KB A tile from shared memory whatever N is, so
  // small-N MMAs are operand-fetch bound), two accumulator stages where TMEM allows:
  //     same-res:  Cout>=256 (1,256)   128 (2,128)   64 (1,64,concat)   <=32 (4,Cout,concat)
  //     up:        Cout>=256 (1,256, 4 groups)   128 (1,128, 2 groups)   64 (2,64, 2 groups)   <=32 (1,Cout,concat)
  // (re-measured after the issue-loop / epilogue rework of this round: the concat product now also wins for Cout = 64
  // same-res and for the Cout <= 32 transposed layers)
  // The search below (cost model) only decides when the preferred configuration is infeasible or leaves most SMs idle.
  int best_r = 0, best_bn = 0, best_cat = 0, best_groups = 1;
  static const int force_groups = [] { const char* e = getenv("MAUA_TC_GROUPS"); return e ? atoi(e) : 0; }();
  const uint32_t a_planes = n_products == 3 ? 2u : 1u;   // n_products: 1 = bf16 hi*hi, 3 = bf16 split, 2 = fp16 (a * (w_hi + w_lo))
  // collapsed transposed conv (see c_shifts): fp16 format, all phases in one item, no concat, N = 4*BN <= 256
  // OFF by default: correct (tests) but measured SLOWER on both layers it applies to — 64->32 @512 up 0.433 vs 0.354 ms,
  // 128->64 @256 up 0.330 vs 0.256 ms, same box, same (R, BN) — although it issues 16 instead of 36 MMAs per K chunk and
  // tile; not profiled yet (suspect: the issue loop of the collapsed form, whose descriptors are not compile-time unrolled
  // per tap, is slower than the tensor pipe).  MAUA_TC_COLL=1 enables it.
  const char* coll_env = getenv("MAUA_TC_COLL");
  const bool coll_on = coll_env && coll_env[0] == '1';
  auto is_coll = [&](int bn, int cat, int groups) {
    return coll_on && up && n_products == 2 && !cat && groups == 1 && bn <= 64;
  };
  auto n_ctas = [&](int r, int bn, int groups) {
    return tiles_x * ceil_div(rows16, (long long)r) * batch * (cout / bn) * groups;
  };
  auto feasible = [&](int r, int bn, int cat, int groups) {
    if (r > rows16 || cout % bn != 0 || bn < 16 || bn > 256 || (bn & (bn - 1)) != 0) return false;
    if (ep.rgb_out && bn != cout) return false;  // fused ToRGB needs every output channel in one CTA
    if (cat && (n_products == 1 || bn > 64)) return false;
    if (!up && groups != 1) return false;
    const int blk = cat ? 2 * bn : bn;  // concat mode doubles the accumulator width
    if (r * blk * (nphase / groups) > tmem_cap) return false;
    const uint32_t plane = align1k((uint32_t)((TH * r + (up ? 1 : 2)) * (up ? TW + 1 : TW + 2) * kc * 2));
    const uint32_t b_st = 2u * bn * kc * 2u * (is_coll(bn, cat, groups) ? 4u : 1u);
    return a_planes * plane + (is_coll(bn, cat, groups) ? 3 : 4) * b_st <= budget;  // A halo + a B ring that hides TMA latency
  };
  auto search = [&]() {
    double best_cost = 1e30;
    long long best_ctas = 0;
    for (int groups = 1; groups <= (up ? 4 : 1); groups <<= 1) {
      if (up && force_groups && groups != force_groups) continue;
      const int nph_item = nphase / groups;  // accumulator phases held by one work item
      for (int r = 4; r >= 1; r >>= 1)
        for (int bn = 256; bn >= 16; bn >>= 1)
          for (int cat = 0; cat <= 1; ++cat) {
            if (!feasible(r, bn, cat, groups)) continue;
            const int blk = cat ? 2 * bn : bn;
            const long long ctas = n_ctas(r, bn, groups);
            // modelled tensor-pipe cycles per MMA of N columns: max(math N/2, operand fetch (4 KB A + 32N B) at
            // ~115 B/clk); a single accumulator stage exposes the epilogue (penalty); L2->SMEM bytes per cycle as a
            // secondary term (the halo is re-loaded once per phase group)
            auto mma_cycles = [](double n) { const double f = (4096.0 + 32.0 * n) / 115.0; return n / 2.0 > f ? n / 2.0 : f; };
            const double per_kstep = n_products == 2 ? (cat ? mma_cycles(2.0 * bn) : 2.0 * mma_cycles(bn))
                                     : cat ? mma_cycles(2.0 * bn) + mma_cycles(bn) : (n_products > 1 ? 3.0 : 1.0) * mma_cycles(bn);
            const double traffic = (groups * (TH * r + 2) * 10.0 / 9.0 + bn) / ((double)r * bn);
            const double cost = per_kstep / bn * (2 * r * blk * nph_item <= 512 ? 1.0 : 1.2) + 0.15 * traffic;
            const bool enough = ctas >= 120, best_enough = best_ctas >= 120;
            const bool better = best_r == 0 || (enough && !best_enough) ||
                                (enough == best_enough && (enough ? cost < best_cost : ctas > best_ctas));
            if (better) { best_r = r; best_bn = bn; best_cat = cat; best_groups = groups; best_cost = cost; best_ctas = ctas; }
          }
    }
  };
  if (f_r) {
    if (feasible(f_r, f_bn, f_cat, f_groups)) { best_r = f_r; best_bn = f_bn; best_cat = f_cat; best_groups = f_groups; }
  } else {
    int pr, pbn, pcat = 0, pg = 1;
    if (cout >= 256) { pr = 1; pbn = 256; pg = up ? 4 : 1; }
    else if (cout == 128) { pr = up ? 1 : 2; pbn = 128; pg = up ? 2 : 1; }
    else if (cout == 64) { pr = up ? 2 : 1; pbn = 64; pg = up ? 2 : 1; pcat = up ? 0 : 1; }
    else { pr = up ? 1 : 4; pbn = cout; pcat = 1; }
    if (n_products == 2 && cout <= 64) {
      // fp16 (tools/tune_tc2.py --prod 2, batch 8): same-res layers take the concat MMA (one N = 2*BN MMA per K-step:
      // 64->64 @512 R=2 0.258 ms vs 0.304 without; 32->32 @1024 R=4 0.434 vs 0.556); the transposed layers are paced by
      // their epilogue's HBM stores, where the concat's second TMEM load + add per chunk costs more than the MMAs it
      // saves (128->64 @256 up: R=2, 2 groups, no concat 0.228 ms vs 0.259; 64->32 @512 up: R=2 no concat 0.301 vs 0.335)
      pcat = up ? 0 : 1;
      pr = 2;
      if (cout == 64) pg = up ? 2 : 1;
      else { pr = up ? 2 : 4; pg = 1; }
      // collapsed taps (c_shifts): all four phases in one item; two accumulator stages need 4*BN*R*2 <= 512 columns
      if (up && coll_on) { pg = 1; pr = cout == 64 ? 1 : 2; }
    }
    if (force_groups && up) pg = force_groups;
    while (pr > 1 && pr > rows16) pr >>= 1;
    // keep (most of) the 148 SMs busy: first fewer stacked tiles, then more phase groups, then narrower N
    while (feasible(pr, pbn, pcat, pg) && n_ctas(pr, pbn, pg) < 120) {
      if (pr > 1) pr >>= 1;
      else if (up && pg < 4 && !force_groups) pg <<= 1;
      else if (pbn > 64 && !ep.rgb_out) pbn >>= 1;
      else break;
    }
    // Wave quantisation of the persistent grid: 512->512 @64^2 at batch 8 is 512 (R=1, BN=256) items = 3.46 waves of the
    // 148 SMs (74 CTA pairs), i.e. a 4th wave at 46 % occupancy.  Half-width N tiles double the item count (6.9 -> 7
    // waves); measured with CTA pairs: 0.280 ms vs 0.294 ms.  Only taken when it recovers >= 8 % of the grid.
    if (!up && n_products == 3 && pbn == 256 && !ep.rgb_out && feasible(pr, 128, pcat, pg)) {
      const double slots = device_sm_count();
      auto wave_eff = [&](int bn) {
        const double it = (double)n_ctas(pr, bn, pg);
        return it / (std::ceil(it / slots) * slots);
      };
      if (wave_eff(256) < 0.9 && wave_eff(128) > wave_eff(256) + 0.08) pbn = 128;
    }
    if (feasible(pr, pbn, pcat, pg)) { best_r = pr; best_bn = pbn; best_cat = pcat; best_groups = pg; }
    else search();
  }
  const int unsupported = f_r ? MAUA_E_ARG : MAUA_E_UNSUPPORTED;  // a forced configuration must not fall back to v1
  if (f_r) set_error("modconv_tc(v2): forced configuration %d,%d,%d,%d is infeasible", f_r, f_bn, f_cat, f_groups);
  if (best_r == 0) return unsupported;
  const int R = best_r, bn = best_bn;
  p.R = R;
  p.BN = bn;
  p.n_tiles = cout / bn;
  p.HW_ = up ? TW + 1 : TW + 2;
  p.HH_ = TH * R + (up ? 1 : 2);
  p.tiles_x = (int)tiles_x;
  p.tiles_y = (int)ceil_div(rows16, (long long)R);
  p.n_kchunks = n_kchunks;
  p.a_planes = (int)a_planes;
  p.b_planes = n_products == 1 ? 1 : 2;
  p.a_plane = align1k((uint32_t)(p.HW_ * p.HH_ * kc * 2));
  p.cat = best_cat;
  const int blk_cols = p.cat ? 2 * bn : bn;
  p.n_groups = best_groups;
  p.n_phase = nphase / best_groups;
  int cols = 32;
  // CTA pairs (cta_group::2) for the split-bf16 layers with wide N tiles: those are bound by shared-memory bandwidth
  // (see the kernel's PAIR note).  MAUA_TC_PAIR=0 disables, =1 enables wherever the mode allows (read per call).
  const char* pair_env = getenv("MAUA_TC_PAIR");
  const int pair_req = pair_env ? atoi(pair_env) : MAUA_TC_PAIR_DEFAULT;
  const long long ptiles = tiles_x * p.tiles_y * batch;
  const bool pair = pair_req != 0 && n_products == 3 && !p.cat && bn >= (pair_req == 1 ? 128 : 32) && ptiles >= 2 &&
                    ptiles < (1LL << 30);
  p.pair = pair ? 1 : 0;
  p.n_ptiles = (int)(ptiles < (1LL << 30) ? ptiles : 0);
  const int bn_cta = pair ? bn / 2 : bn;
  p.coll = is_coll(bn, p.cat, p.n_groups) ? 1 : 0;
  const uint32_t a_stage = a_planes * p.a_plane, b_stage = 2u * bn_cta * kc * 2u * (p.coll ? 4u : 1u);
  p.SA = (2 * a_stage + 4 * b_stage <= budget) ? 2 : 1;
  int sb = (int)((budget - (uint32_t)p.SA * a_stage) / b_stage);
  // small layers: keep ALL weight tiles of the layer in shared memory for the lifetime of the persistent CTA
  p.resident_b = (!p.coll && p.n_tiles == 1 && p.n_groups == 1 && 9 * n_kchunks <= sb && 9 * n_kchunks <= 36) ? 1 : 0;
  if (p.resident_b) sb = 9 * n_kchunks;
  else if (sb > 12) sb = 12;
  if (sb < 2) return unsupported;
  p.SB = sb;
  const size_t smem = (size_t)p.SA * a_stage + (size_t)p.SB * b_stage + 8 * (2 * p.SA + 2 * p.SB + 6) + 1024 +
                      6 * (size_t)cout * 4 + 32;
  if (smem > 227 * 1024) return unsupported;
  const long long items = (pair ? (ptiles + 1) / 2 : ptiles) * p.n_tiles * p.n_groups;
  if (items >= (1LL << 31)) return MAUA_E_UNSUPPORTED;
  p.n_items = (int)items;
  p.per_group = (int)(items / p.n_groups);
  {
    const uint32_t ds[4] = {(uint32_t)p.per_group, (uint32_t)p.n_tiles, (uint32_t)p.tiles_x, (uint32_t)p.tiles_y};
    for (int i = 0; i < 4; ++i) {
      uint32_t sh = 0;
      while ((1ull << sh) < ds[i]) ++sh;
      p.fd_s[i] = sh;
      p.fd_m[i] = (uint32_t)((((1ull << sh) - ds[i]) << 32) / ds[i] + 1);
    }
  }
  p.AS = (2 * blk_cols * R * p.n_phase <= 512) ? 2 : 1;
  static const int dbg = [] { const char* e = getenv("MAUA_TC_DBG"); return e ? atoi(e) : 0; }();
  p.dbg = dbg;
  cols = 32;
  while (cols < p.AS * blk_cols * R * p.n_phase) cols <<= 1;
  p.tmem_cols = (uint32_t)cols;
  const int n_sm = device_sm_count();
  const long long grid = pair ? 2 * (items < n_sm / 2 ? items : n_sm / 2) : (items < n_sm ? items : n_sm);
  static const bool debug = [] { const char* e = getenv("MAUA_TC_DEBUG"); return e && e[0] == '1'; }();
  if (debug)
    fprintf(stderr, "[modconv_tc2] %s B%d %d->%d @%dx%d: R=%d BN=%d cat=%d groups=%d resB=%d AS=%d SA=%d SB=%d smem=%zuKB tmem=%u items=%lld grid=%lld\n",
            up ? "up" : "same", batch, cin, cout, h, w, R, bn, p.cat, p.n_groups, p.resident_b, p.AS, p.SA, p.SB, smem / 1024, p.tmem_cols, items, grid);

  set_conv_config("v2 up=%d R=%d BN=%d cat=%d groups=%d prod=%d resB=%d AS=%d SA=%d SB=%d items=%lld grid=%lld coll=%d pair=%d", up,
                  R, bn, p.cat, p.n_groups, n_products, p.resident_b, p.AS, p.SA, p.SB, items, grid, p.coll, p.pair);
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  const auto swz = kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  const cuuint64_t adims[4] = {(cuuint64_t)cin, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)batch};
  const cuuint64_t astr[3] = {(cuuint64_t)cin * 2, (cuuint64_t)w * cin * 2, (cuuint64_t)h * w * cin * 2};
  const cuuint32_t abox[4] = {(cuuint32_t)kc, (cuuint32_t)p.HW_, (cuuint32_t)p.HH_, 1};
  const cuuint64_t bdims[3] = {(cuuint64_t)cin, (cuuint64_t)cout, 9};
  const cuuint64_t bstr[2] = {(cuuint64_t)cin * 2, (cuuint64_t)cout * cin * 2};
  const cuuint32_t bbox[3] = {(cuuint32_t)kc, (cuuint32_t)bn_cta, 1};   // (pairs: each CTA loads half of the N tile's rows)
  int rc;
  const auto dt = n_products == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  if ((rc = tmap::encode(&ta_hi, dt, x_hi, 4, adims, astr, abox, swz))) return rc;
  if ((rc = tmap::encode(&tb_hi, dt, w_hi, 3, bdims, bstr, bbox, swz))) return rc;
  if (a_planes > 1) {
    if ((rc = tmap::encode(&ta_lo, dt, x_lo, 4, adims, astr, abox, swz))) return rc;
  } else {
    ta_lo = ta_hi;
  }
  if (p.b_planes > 1) {
    if ((rc = tmap::encode(&tb_lo, dt, w_lo, 3, bdims, bstr, bbox, swz))) return rc;
  } else {
    tb_lo = tb_hi;
  }
#define MAUA_TC2_LAUNCH(KCV, UPV, MODEV)                                                                            \
  do {                                                                                                              \
    MAUA_CHECK_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(modconv_tc2_kernel<KCV, UPV, MODEV, false>), smem)); \
    MAUA_CHECK_CUDA(launch_chain(modconv_tc2_kernel<KCV, UPV, MODEV, false>, dim3((unsigned)grid), dim3(THREADS), smem, st, 1, \
                                 ta_hi, ta_lo, tb_hi, tb_lo, p, ep));                                                  \
  } while (0)
  const int mode = n_products == 1 ? 0 : (n_products == 2 ? (p.cat ? 4 : 3) : (p.cat ? 2 : 1));
#define MAUA_TC2_LAUNCH_PAIR(KCV, UPV, MODEV)                                                                       \
  do {                                                                                                              \
    auto kern = modconv_tc2_kernel<KCV, UPV, MODEV, true>;                                                          \
    MAUA_CHECK_CUDA(ensure_dyn_smem(reinterpret_cast<const void*>(kern), smem));                                    \
    MAUA_CHECK_CUDA(launch_chain(kern, dim3((unsigned)grid), dim3(THREADS), smem, st, 2, ta_hi, ta_lo, tb_hi, tb_lo, p, ep)); \
  } while (0)
#define MAUA_TC2_MODES(UPV)                                                   \
  switch (mode) {                                                             \
    case 0: MAUA_TC2_LAUNCH(32, UPV, 0); break;                               \
    case 1: if (pair) MAUA_TC2_LAUNCH_PAIR(32, UPV, 1); else MAUA_TC2_LAUNCH(32, UPV, 1); break; \
    case 2: MAUA_TC2_LAUNCH(32, UPV, 2); break;                               \
    case 3: MAUA_TC2_LAUNCH(32, UPV, 3); break;                               \
    default: MAUA_TC2_LAUNCH(32, UPV, 4); break;                              \
  }
  if (up) { MAUA_TC2_MODES(true) } else { MAUA_TC2_MODES(false) }
#undef MAUA_TC2_MODES
#undef MAUA_TC2_LAUNCH_PAIR
#undef MAUA_TC2_LAUNCH
  MAUA_CHECK_LAUNCH("modconv_tc(v2)");
  return MAUA_OK;
}

}  // namespace maua
