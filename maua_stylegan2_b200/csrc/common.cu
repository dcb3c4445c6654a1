// Error state, launch counter and ABI version of libmaua_b200.so.
#include "common.cuh"

#include <atomic>
#include <cstdarg>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>

namespace maua {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static thread_local char g_conv_cfg[256] = "";

void set_conv_config(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_conv_cfg, sizeof(g_conv_cfg), fmt, ap);
  va_end(ap);
}

static std::mutex g_dev_mu;
static std::map<int, int> g_sm_count;                              // device -> SM count
static std::map<std::pair<const void*, int>, size_t> g_dyn_smem;   // (kernel, device) -> limit already set

bool pdl_enabled() {   // read per call: tests and A/B runs toggle it
  const char* e = getenv("MAUA_PDL");
  return !(e && e[0] == '0');
}

int device_sm_count() {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_dev_mu);
  auto it = g_sm_count.find(dev);
  if (it != g_sm_count.end()) return it->second;
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  g_sm_count[dev] = n;
  return n;
}

cudaError_t ensure_dyn_smem(const void* kernel, size_t bytes) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_dev_mu);
  size_t& have = g_dyn_smem[std::make_pair(kernel, dev)];
  if (bytes <= have) return cudaSuccess;  // (no runtime call on the hot path, nor during CUDA-graph capture after warm-up)
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) have = bytes;
  return e;
}

}  // namespace maua

extern "C" {

int maua_abi_version(void) { return MAUA_ABI_VERSION; }

const char* maua_last_error(void) { return maua::g_err; }

long long maua_launch_count(void) { return maua::g_launches.load(std::memory_order_relaxed); }

const char* maua_modconv_tc_last_config(void) { return maua::g_conv_cfg; }

}  // extern "C"
