// Internal helpers shared by the translation units of libtribe_b200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tribe {
int set_error(int code, const char* msg);
int set_cuda_error(cudaError_t e, const char* where);
void count_launch();

inline int grid_for(int64_t work_items, int per_block, int max_blocks) {
  int64_t b = (work_items + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return static_cast<int>(b);
}
}  // namespace tribe

#define TRIBE_CHECK_LAUNCH(where)                                  \
  do {                                                             \
    ::tribe::count_launch();                                       \
    cudaError_t e__ = cudaGetLastError();                          \
    if (e__ != cudaSuccess) return ::tribe::set_cuda_error(e__, where); \
  } while (0)
