// Error reporting and launch accounting for the C ABI (include/tribe_b200.h).
#include <atomic>
#include <stdio.h>
#include <string.h>

#include "tribe_b200.h"
#include "tribe_internal.h"

namespace tribe {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg);
  return code;
}
int set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
  return static_cast<int>(e);
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace tribe

extern "C" const char* tribe_last_error(void) { return tribe::g_err; }
extern "C" int tribe_abi_version(void) { return 1; }
extern "C" int64_t tribe_launch_count(void) { return tribe::g_launches.load(); }
