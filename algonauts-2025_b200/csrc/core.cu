// Error reporting and launch accounting for the C ABI (include/tribe_b200.h).
#include <atomic>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
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
static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("TRIBE_PDL");
    v = (e && atoi(e) != 0) ? 1 : 0;  // opt-in: measured neutral on B200 inside whole-step CUDA graphs (profiles/r02_pdl_ab.txt)
    g_pdl.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}
}  // namespace tribe

extern "C" const char* tribe_last_error(void) { return tribe::g_err; }
extern "C" int tribe_abi_version(void) { return 1; }
extern "C" int tribe_peek_last_error(void) { return static_cast<int>(cudaPeekAtLastError()); }
extern "C" int tribe_take_last_error(void) { return static_cast<int>(cudaGetLastError()); }
extern "C" int tribe_set_pdl(int32_t on) {
  tribe::g_pdl.store(on ? 1 : 0, std::memory_order_relaxed);
  return TRIBE_OK;
}
extern "C" int64_t tribe_launch_count(void) { return tribe::g_launches.load(); }
