// Error state, launch counter and ABI version of libmaua_b200.so.
#include "common.cuh"

#include <atomic>
#include <cstdarg>
#include <cstring>

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

}  // namespace maua

extern "C" {

int maua_abi_version(void) { return MAUA_ABI_VERSION; }

const char* maua_last_error(void) { return maua::g_err; }

long long maua_launch_count(void) { return maua::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
