// Experiment: can a K-major swizzled UMMA A operand start at an arbitrary ROW of a TMA-written tile?
// (needed to reuse ONE halo tile for the 9 taps of a 3x3 conv instead of 9 shifted TMA loads)
//   A_all: [R rows][KC] bf16 written by one TMA box with SWIZZLE_128B (KC=64) or SWIZZLE_64B (KC=32)
//   D[i][n] = sum_k A_all[s+i][k] * B[n][k],  i < 128, n < 32, for several row shifts s and base_offset encodings.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I../include -o exp_shifted_desc exp_shifted_desc.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>

#include "../maua_stylegan2_b200/csrc/common.cuh"
#include "../maua_stylegan2_b200/csrc/sm100_ptx.cuh"
#include "../maua_stylegan2_b200/csrc/tmap.cuh"

namespace maua { void set_error(const char* fmt, ...) {} void count_launch(int) {} }

using namespace maua;
using namespace maua::ptx;

template <int KC>
__global__ void __launch_bounds__(128, 1) exp_kernel(const __grid_constant__ CUtensorMap tm_a,
                                                     const __grid_constant__ CUtensorMap tm_b, float* out, int R,
                                                     int shift, int bo_mode, int pitch_rows) {
  constexpr uint32_t ROW = KC * 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_smem = base;
  const uint32_t b_smem = base + ((R * ROW + 1023u) & ~1023u);
  const uint32_t bar = b_smem + 32 * ROW + 1024;
  const uint32_t bar2 = bar + 8;
  const uint32_t slot = bar + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(bar2, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(slot, 32);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, R * ROW + 32 * ROW);
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(a_smem), "l"(reinterpret_cast<uint64_t>(&tm_a)), "r"(bar), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(b_smem), "l"(reinterpret_cast<uint64_t>(&tm_b)), "r"(bar), "r"(0), "r"(0) : "memory");
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t a_start = a_smem + shift * ROW;
    uint64_t da = make_kmajor_desc(a_start, ROW);
    da &= ~(0x3FFFull << 32);
    da |= (uint64_t)(((uint32_t)pitch_rows * ROW) >> 4) << 32;  // SBO = pitch between 8-row groups
    uint64_t bo = 0;
    if (bo_mode == 1) bo = (a_start >> 7) & 7;
    if (bo_mode == 2) bo = (a_start / ROW) & 7;  // row index mod 8
    da |= bo << 49;
    const uint64_t db = make_kmajor_desc(b_smem, ROW);
    const uint32_t idesc = make_idesc_bf16(128, 32);
    for (int k = 0; k < KC / 16; ++k) umma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, k > 0);
    umma_commit(bar2);
  }
  mbar_wait(bar2, 0);
  tc_fence_after();
  uint32_t r[16];
  for (int c = 0; c < 32; c += 16) {
    tmem_ld_x16(tmem + ((uint32_t)(warp * 32) << 16) + c, r);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 32 + c + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

template <int KC>
int run(int R, int pitch) {
  std::vector<__nv_bfloat16> A(R * KC), B(32 * KC);
  std::vector<float> Af(R * KC), Bf(32 * KC);
  srand(1);
  for (int i = 0; i < R * KC; ++i) { float v = (rand() % 2001 - 1000) / 500.f; A[i] = __float2bfloat16_rn(v); Af[i] = bf(v); }
  for (int i = 0; i < 32 * KC; ++i) { float v = (rand() % 2001 - 1000) / 500.f; B[i] = __float2bfloat16_rn(v); Bf[i] = bf(v); }
  __nv_bfloat16 *dA, *dB; float* dO;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dO, 128 * 32 * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ta, tb;
  cuuint64_t da[2] = {(cuuint64_t)KC, (cuuint64_t)R}, sa[1] = {(cuuint64_t)KC * 2};
  cuuint32_t ba[2] = {(cuuint32_t)KC, (cuuint32_t)R};
  cuuint64_t db[2] = {(cuuint64_t)KC, 32};
  cuuint32_t bb[2] = {(cuuint32_t)KC, 32};
  auto swz = KC == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  if (R > 256) { printf("R too big for one box\n"); return 1; }
  if (tmap::encode(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dA, 2, da, sa, ba, swz)) { printf("encode A failed\n"); return 1; }
  if (tmap::encode(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, dB, 2, db, sa, bb, swz)) { printf("encode B failed\n"); return 1; }
  const size_t smem = R * KC * 2 + 32 * KC * 2 + 4096 + 1024;
  cudaFuncSetAttribute(exp_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  std::vector<float> O(128 * 32);
  const int shifts[] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 16, 17, 65, 100};
  for (int mode = 0; mode < 3; ++mode) {
    if (mode != 0) continue;
    printf("KC=%d pitch %d:", KC, pitch);
    for (int s : shifts) {
      if (s + 15 * pitch + 8 > R) continue;
      cudaMemset(dO, 0, 128 * 32 * 4);
      exp_kernel<KC><<<1, 128, smem>>>(ta, tb, dO, R, s, mode, pitch);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf(" [s=%d CUDA error %s]", s, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int i = 0; i < 128; ++i)
        for (int n = 0; n < 32; ++n) {
          double ref = 0;
          for (int k = 0; k < KC; ++k) ref += (double)Af[(s + (i / 8) * pitch + (i % 8)) * KC + k] * Bf[n * KC + k];
          maxerr = fmax(maxerr, fabs(ref - O[i * 32 + n]));
        }
      printf(" s=%d:%s(%.2g)", s, maxerr < 1e-3 ? "OK" : "BAD", maxerr);
    }
    printf("\n");
  }
  return 0;
}

int main() {
  int rc = 0;
  for (int pitch : {8, 9, 10, 17}) { rc |= run<64>(232, pitch); rc |= run<32>(232, pitch); }
  return rc;
}
