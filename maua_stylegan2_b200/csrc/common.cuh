// Shared host/device helpers for libmaua_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/maua_b200.h"

namespace maua {

// thread-local error text + process-wide launch counter (defined in common.cu)
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
void set_conv_config(const char* fmt, ...);  // what maua_modconv_tc_last_config() reports (thread-local)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Per-DEVICE caches (defined in common.cu).  cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device/context
// attribute and the SM count drives the tile policies: a process that renders on cuda:0 and then on cuda:1 must set /
// query them again on the second device (a process-wide `static bool done` silently broke every >48 KB launch there).
int device_sm_count();                                   // SMs of the CURRENT device (148 on B200)
cudaError_t ensure_dyn_smem(const void* kernel, size_t bytes);  // raise the kernel's dynamic-smem limit to >= bytes

#define MAUA_CHECK_ARG(cond, ...)            \
  do {                                       \
    if (!(cond)) {                           \
      ::maua::set_error(__VA_ARGS__);        \
      return MAUA_E_ARG;                     \
    }                                        \
  } while (0)

// Checks the launch (not the execution: everything is asynchronous on the caller's stream).
#define MAUA_CHECK_LAUNCH(name)                                                       \
  do {                                                                                \
    cudaError_t _e = cudaGetLastError();                                              \
    if (_e != cudaSuccess) {                                                          \
      ::maua::set_error("%s: launch failed: %s", name, cudaGetErrorString(_e));       \
      return MAUA_E_CUDA;                                                             \
    }                                                                                 \
    ::maua::count_launch();                                                           \
  } while (0)

#define MAUA_CHECK_CUDA(expr)                                                         \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      ::maua::set_error("%s failed: %s", #expr, cudaGetErrorString(_e));              \
      return MAUA_E_CUDA;                                                             \
    }                                                                                 \
  } while (0)

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) {
  return (a + b - 1) / b;
}

__host__ __device__ __forceinline__ int floor_div(int a, int b) {
  int q = a / b;
  return (q * b > a) ? q - 1 : q;
}

// fp32 -> (hi, lo) bf16 pair with hi + lo == x up to 2^-17 relative.
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// fp32 pair -> packed fp16x2 (low half = a), round-to-nearest, SATURATING to +-65504 instead of overflowing to inf:
// the "f16" activation format of the tensor-core path must stay finite whatever a checkpoint's dynamic range is.
__device__ __forceinline__ uint32_t pack_f16x2_sat(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %2, %1;" : "=r"(r) : "f"(a), "f"(b));  // d = {hi: first source, lo: second}
  return r;
}
__device__ __forceinline__ __half f16_sat(float a) {
  const uint32_t r = pack_f16x2_sat(a, 0.f);
  return __ushort_as_half((unsigned short)(r & 0xFFFFu));
}
// fp32 -> (hi, lo) fp16 pair with hi + lo == x up to 2^-22 relative (weights of the "f16" path)
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  hi = f16_sat(x);
  lo = f16_sat(x - __half2float(hi));
}

// Packed dual fp32 FMA (sm_100 `fma.rn.f32x2`): two IEEE-rounded FMAs per issue slot; each lane rounds exactly like
// a scalar fmaf, so it does not change results — it halves the FMA instruction count of the FIR kernels.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}

__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}

// 256-bit global store (sm_100: STG.E.ENL2.256): 8 consecutive floats, address 32-byte aligned
__device__ __forceinline__ void st_global_v8(float* p, float2 a, float2 b, float2 c, float2 d) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x),
               "f"(c.y), "f"(d.x), "f"(d.y)
               : "memory");
}
__device__ __forceinline__ void st_global_v8_b32(void* p, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4,
                                                 uint32_t a5, uint32_t a6, uint32_t a7) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5),
               "r"(a6), "r"(a7)
               : "memory");
}

// Programmatic dependent launch for the kernels of the synthesis chain (conv -> blur_act -> conv ...): the next kernel's
// launch, CTA scheduling and global-memory-free setup overlap the tail of the running one (sm100_ptx.cuh: pdl_wait).
// MAUA_PDL=0 launches them with plain stream order.  `cluster_x` > 1 adds the cluster dimension (CTA pairs).
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chain(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x,
                                Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = (unsigned)cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Exact x / d for 0 <= x < 2^31 and a runtime-constant d >= 1 without the ~25-instruction integer-division sequence:
// q = (umulhi(x, m) + x) >> s with m = floor(2^32 * (2^s - d) / d) + 1, s = ceil(log2 d)  (round-up method, 33-bit magic).
struct FastDiv { uint32_t m, s; };
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  uint32_t sh = 0;
  while ((1ull << sh) < d) ++sh;
  f.s = sh;
  f.m = (uint32_t)((((1ull << sh) - d) << 32) / d + 1);
  return f;
}
__device__ __forceinline__ int fast_div(int x, FastDiv f) {
  return (int)((__umulhi((uint32_t)x, f.m) + (uint32_t)x) >> f.s);
}

__device__ __forceinline__ float lrelu_scaled(float v, float slope, float scale) {
  return __fmul_rn(v > 0.f ? v : __fmul_rn(v, slope), scale);
}

}  // namespace maua
