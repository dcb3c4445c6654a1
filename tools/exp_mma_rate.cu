// Experiment: tcgen05.mma issue/throughput per shared-memory operand layout (timing only; operand values are garbage).
// One CTA per SM; one elected lane issues GROUPS x {MMAs of 128 x N x 16, kind::f16} back to back, commits, waits.
//   layout 0: SWIZZLE_64B, 64-byte rows (what modconv_tc2 v2 uses; SBO = halo pitch 10 px * 64 B)
//   layout 1: SWIZZLE_128B, 128-byte rows (CUTLASS default; SBO = 1024)
//   layout 2: no swizzle, "interleaved" K-chunk-major planes: 8 rows x 16 B core matrices contiguous,
//             SBO = 10 px * 16 B, LBO = plane stride
// Prints cycles per MMA for N in {32, 64, 128, 256}.   nvcc -arch=sm_100a -o tools/bin/exp_mma_rate tools/exp_mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../maua_stylegan2_b200/csrc/sm100_ptx.cuh"
using namespace maua::ptx;

// flags bit0: A start address walks the 9 tap offsets of a 10-pixel-pitch halo; bit1: 12 other warps hammer TMEM with
// tcgen05.ld (the conv epilogue's access pattern) while the MMAs run; bit2: "concat" mix: N, N, N/2, N/2 per group
__global__ void __launch_bounds__(448, 1) k(int layout, int n, int groups, long long* out, int flags) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 200 * 1024, slot = bar + 16;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); *reinterpret_cast<volatile int*>(smem_raw + 201 * 1024) = 0; }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (warp == 0) {
    const bool leader = elect_one_sync();
    const uint32_t idesc = make_idesc_bf16(128, (uint32_t)n);
    uint64_t da, db, kstep;
    const uint32_t a_base = base, b_base = base + 96 * 1024;
    if (layout == 0) {
      da = (make_kmajor_desc(a_base, 64) & ~(0x3FFFull << 32)) | ((uint64_t)((10 * 64) >> 4) << 32);
      db = make_kmajor_desc(b_base, 64);
      kstep = 2;  // +32 B inside the 64-byte row
    } else if (layout == 1) {
      da = make_kmajor_desc(a_base, 128);
      db = make_kmajor_desc(b_base, 128);
      kstep = 2;
    } else {
      const uint64_t a_lbo = (66 * 10 * 16) >> 4, b_lbo = ((uint64_t)n * 16) >> 4;
      da = (uint64_t)((a_base >> 4) & 0x3FFF) | (a_lbo << 16) | ((uint64_t)((10 * 16) >> 4) << 32) | (1ull << 46);
      db = (uint64_t)((b_base >> 4) & 0x3FFF) | (b_lbo << 16) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
      kstep = 0;  // K step advances by 2 planes: handled below
    }
    const uint64_t a_k = layout == 2 ? 2 * ((66ull * 10 * 16) >> 4) : kstep;
    const uint64_t b_k = layout == 2 ? 2 * (((uint64_t)n * 16) >> 4) : kstep;
    __syncwarp();
    const long long t0 = clock64();
    const uint32_t idesc_h = make_idesc_bf16(128, (uint32_t)(n / 2));
    const uint32_t row16 = layout == 1 ? 8 : (layout == 0 ? 4 : 1);  // one pixel row of the operand in 16-byte units
    for (int g = 0; g < groups; ++g) {
      const int tap = (flags & 1) ? g % 9 : 0;
      const uint64_t dat = da + (uint64_t)(((tap / 3) * 10 + tap % 3) * row16);
      if (leader) {
        if (flags & 4) {
          umma_bf16(tmem, dat, db, idesc, g > 0);
          umma_bf16(tmem, dat + a_k, db + b_k, idesc, 1u);
          umma_bf16(tmem, dat + 64, db, idesc_h, 1u);
          umma_bf16(tmem, dat + 64 + a_k, db + b_k, idesc_h, 1u);
        } else {
          // one "tap": two K=16 steps on two accumulators (like R = 2 stacked tiles)
          umma_bf16(tmem, dat, db, idesc, g > 0);
          umma_bf16(tmem, dat + a_k, db + b_k, idesc, 1u);
          umma_bf16(tmem + 256, dat + 40 * row16 / 4, db, idesc, g > 0);
          umma_bf16(tmem + 256, dat + 40 * row16 / 4 + a_k, db + b_k, idesc, 1u);
        }
      }
    }
    if (leader) umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    if (leader && blockIdx.x == 0) out[0] = t1 - t0;
    if (leader) *reinterpret_cast<volatile int*>(smem_raw + 201 * 1024) = 1;  // stop the LDTM warps
  } else if (warp >= 2 && (flags & 2)) {
    // epilogue-like TMEM readers: warp w reads lane quadrant (w & 3), sweeping 16-column chunks
    const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t sink = 0, col = (uint32_t)(warp >> 2) * 16;
    while (*reinterpret_cast<volatile int*>(smem_raw + 201 * 1024) == 0) {
      uint32_t r[16];
      tmem_ld_x16(lane_addr + col, r);
      tmem_ld_wait();
      sink += r[0] + r[15];
      col = (col + 48) & 511;
    }
    if (sink == 0x12345678u) out[1] = sink;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  const size_t smem = 202 * 1024 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int groups = 2000;
  const char* names[3] = {"SWIZZLE_64B/64B rows ", "SWIZZLE_128B/128B rows", "no swizzle interleaved"};
  for (int flags : {0, 1, 2, 3, 4, 6, 7})
  for (int layout = 0; layout < 3; ++layout)
    for (int n : {32, 64, 128, 256}) {
      if (flags && (layout != 0 || n > 128)) continue;
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        k<<<148, 448, smem>>>(layout, n, groups, d, flags);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("layout %d n %d: %s\n", layout, n, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      }
      printf("flags %d  %s  N=%3d : %7.1f cycles / MMA   (math floor %d)\n", flags, names[layout], n, (double)h / (4.0 * groups), n / 2);
    }
  return 0;
}
