// Internal helpers shared by the translation units of libtribe_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tribe {
int set_error(int code, const char* msg);
int set_cuda_error(cudaError_t e, const char* where);
void count_launch();

// Programmatic dependent launch (PDL).  Kernels that start with TRIBE_PDL_ENTRY() may be launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: `griddepcontrol.launch_dependents` at their top lets the NEXT such
// kernel's CTAs be scheduled (and run their prologue: barrier init, TMEM allocation, tensor-map prefetch) as soon as SM
// resources free up, while `griddepcontrol.wait` holds every CTA until the previous kernel has completed and its
// memory is visible — so ordering is exactly stream order, only launch latency and prologues overlap the predecessor's
// tail.  A kernel launched with the attribute MUST execute the wait in every CTA; launch_k() is therefore only used
// for kernels that contain the macro.  Opt-in (TRIBE_PDL=1 / tribe_set_pdl): without it launch_k is a plain launch.
bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg;
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

inline int grid_for(int64_t work_items, int per_block, int max_blocks) {
  int64_t b = (work_items + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return static_cast<int>(b);
}
}  // namespace tribe

#define TRIBE_PDL_ENTRY()                                        \
  do {                                                           \
    asm volatile("griddepcontrol.launch_dependents;");           \
    asm volatile("griddepcontrol.wait;" ::: "memory");           \
  } while (0)

#define TRIBE_CHECK_LAUNCH(where)                                  \
  do {                                                             \
    ::tribe::count_launch();                                       \
    cudaError_t e__ = cudaGetLastError();                          \
    if (e__ != cudaSuccess) return ::tribe::set_cuda_error(e__, where); \
  } while (0)
