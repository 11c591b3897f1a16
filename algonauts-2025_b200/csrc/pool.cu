// nn.AdaptiveAvgPool1d over the last dimension as a streaming HBM kernel (reference: algonauts2025/model.py:60,119-122;
// (B, 1000, 298) fp32 -> (B, 1000, 100): windows [floor(i*T/T'), ceil((i+1)*T/T')), 96 of length 4 and 4 of length 3).
//
// Rows are contiguous in memory, so a chunk of R rows is ONE contiguous stream: persistent blocks walk the chunks with
// a two-stage cp.async (LDGSTS, 16 B) pipeline into shared memory — the loads of chunk i+1 are in flight while chunk
// i is reduced through a per-block window table and written back as one contiguous, coalesced stream.
// Algorithmic bytes: (t_in + t_out) * 4 per row = 1.592 MB per window.
#include <cuda_runtime.h>
#include <stdint.h>

#include "tribe_b200.h"
#include "tribe_internal.h"

namespace tribe {

__device__ __forceinline__ int pool_win_start(int i, int t_in, int t_out) { return static_cast<int>((static_cast<int64_t>(i) * t_in) / t_out); }
__device__ __forceinline__ int pool_win_end(int i, int t_in, int t_out) {
  return static_cast<int>((static_cast<int64_t>(i + 1) * t_in + t_out - 1) / t_out);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Stage `n` contiguous floats (16-byte aligned base) into shared memory asynchronously; scalar tail by plain loads.
__device__ __forceinline__ void stage_chunk(float* buf, const float* src, int n) {
  const int nv = n >> 2;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) cp_async16(buf + 4 * i, src + 4 * i);
  for (int i = (nv << 2) + threadIdx.x; i < n; i += blockDim.x) buf[i] = __ldg(src + i);
}

constexpr int kPoolRows = 16;  // rows per chunk: 16 * t_in * 4 B is always a multiple of 16 B
__device__ __forceinline__ int rows_in_chunk(int64_t rows, int64_t chunk) {
  const int64_t left = rows - chunk * kPoolRows;
  return static_cast<int>(left < kPoolRows ? left : kPoolRows);
}

__global__ void __launch_bounds__(256) pool_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t rows, int t_in, int t_out) {
  extern __shared__ __align__(16) float sm[];
  const int chunk_floats = kPoolRows * t_in;
  float* buf[2] = {sm, sm + chunk_floats};
  int* wstart = reinterpret_cast<int*>(sm + 2 * chunk_floats);
  int* wlen = wstart + t_out;
  for (int i = threadIdx.x; i < t_out; i += blockDim.x) {
    const int s0 = pool_win_start(i, t_in, t_out);
    wstart[i] = s0;
    wlen[i] = pool_win_end(i, t_in, t_out) - s0;
  }
  const int64_t nchunks = (rows + kPoolRows - 1) / kPoolRows;
  int64_t chunk = blockIdx.x;
  if (chunk < nchunks) stage_chunk(buf[0], x + chunk * chunk_floats, rows_in_chunk(rows, chunk) * t_in);
  cp_async_commit();
  const int ix = threadIdx.x & 127, iy = threadIdx.x >> 7;
  int cur = 0;
  for (; chunk < nchunks; chunk += gridDim.x) {
    const int64_t next = chunk + gridDim.x;
    if (next < nchunks) stage_chunk(buf[cur ^ 1], x + next * chunk_floats, rows_in_chunk(rows, next) * t_in);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const int nr = rows_in_chunk(rows, chunk);
    const float* data = buf[cur];
    float* dst = y + chunk * kPoolRows * t_out;
    // thread = (output index i, row group): the window of i is looked up once and reused for every row of the group;
    // for a fixed row consecutive threads write consecutive outputs (coalesced) and read smem at a stride of ~T/T'.
    for (int i = ix; i < t_out; i += 128) {
      const int s0 = wstart[i], len = wlen[i];
      const float inv = 1.0f / static_cast<float>(len);
      if (len <= 4) {
        const int k1 = len > 1 ? 1 : 0, k2 = len > 2 ? 2 : 0, k3 = len > 3 ? 3 : 0;
        const float m1 = len > 1 ? 1.f : 0.f, m2 = len > 2 ? 1.f : 0.f, m3 = len > 3 ? 1.f : 0.f;
#pragma unroll 4
        for (int r = iy; r < nr; r += 2) {
          const float* row = data + r * t_in + s0;
          dst[r * t_out + i] = (row[0] + m1 * row[k1] + m2 * row[k2] + m3 * row[k3]) * inv;
        }
      } else {
        for (int r = iy; r < nr; r += 2) {
          const float* row = data + r * t_in + s0;
          float acc = 0.f;
          for (int t = 0; t < len; ++t) acc += row[t];
          dst[r * t_out + i] = acc * inv;
        }
      }
    }
    __syncthreads();
    cur ^= 1;
  }
  cp_async_wait<0>();
}

// dx[r, t] = sum over the (contiguous) run of windows containing t of dy[r, i] / len_i.
__global__ void __launch_bounds__(256) pool_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int64_t rows, int t_in, int t_out) {
  extern __shared__ __align__(16) float sm[];
  const int chunk_floats = kPoolRows * t_out;
  float* buf[2] = {sm, sm + chunk_floats};
  float* winv = sm + 2 * chunk_floats;              // t_out : 1 / len_i
  int* ilo = reinterpret_cast<int*>(winv + t_out);  // t_in  : first window containing t
  int* icnt = ilo + t_in;                           // t_in  : number of windows containing t
  for (int i = threadIdx.x; i < t_out; i += blockDim.x) winv[i] = 1.0f / static_cast<float>(pool_win_end(i, t_in, t_out) - pool_win_start(i, t_in, t_out));
  for (int t = threadIdx.x; t < t_in; t += blockDim.x) {
    int lo = static_cast<int>((static_cast<int64_t>(t) * t_out) / t_in);
    int hi = static_cast<int>((static_cast<int64_t>(t + 1) * t_out + t_in - 1) / t_in);
    if (hi > t_out) hi = t_out;
    while (lo < hi && !(t >= pool_win_start(lo, t_in, t_out) && t < pool_win_end(lo, t_in, t_out))) ++lo;
    while (hi > lo && !(t >= pool_win_start(hi - 1, t_in, t_out) && t < pool_win_end(hi - 1, t_in, t_out))) --hi;
    ilo[t] = lo;
    icnt[t] = hi - lo;
  }
  const int64_t nchunks = (rows + kPoolRows - 1) / kPoolRows;
  int64_t chunk = blockIdx.x;
  if (chunk < nchunks) stage_chunk(buf[0], dy + chunk * chunk_floats, rows_in_chunk(rows, chunk) * t_out);
  cp_async_commit();
  const int ix = threadIdx.x & 127, iy = threadIdx.x >> 7;
  int cur = 0;
  for (; chunk < nchunks; chunk += gridDim.x) {
    const int64_t next = chunk + gridDim.x;
    if (next < nchunks) stage_chunk(buf[cur ^ 1], dy + next * chunk_floats, rows_in_chunk(rows, next) * t_out);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const int nr = rows_in_chunk(rows, chunk);
    const float* dyv = buf[cur];
    float* dst = dx + chunk * kPoolRows * t_in;
    for (int t = ix; t < t_in; t += 128) {
      const int lo = ilo[t], cnt = icnt[t];
      if (cnt <= 2) {
        const float w0 = cnt > 0 ? winv[lo] : 0.f, w1 = cnt > 1 ? winv[lo + 1] : 0.f;
        const int k0 = cnt > 0 ? lo : 0, k1 = cnt > 1 ? lo + 1 : k0;
#pragma unroll 4
        for (int r = iy; r < nr; r += 2) {
          const float* row = dyv + r * t_out;
          dst[r * t_in + t] = row[k0] * w0 + row[k1] * w1;
        }
      } else {
        for (int r = iy; r < nr; r += 2) {
          const float* row = dyv + r * t_out;
          float acc = 0.f;
          for (int k = 0; k < cnt; ++k) acc += row[lo + k] * winv[lo + k];
          dst[r * t_in + t] = acc;
        }
      }
    }
    __syncthreads();
    cur ^= 1;
  }
  cp_async_wait<0>();
}

}  // namespace tribe

using namespace tribe;

static int pool_grid(int64_t rows, size_t smem) {
  const int64_t nchunks = (rows + kPoolRows - 1) / kPoolRows;
  int per_sm = static_cast<int>((220 * 1024) / (smem + 1024));
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  const int64_t cap = static_cast<int64_t>(148) * per_sm;
  return static_cast<int>(nchunks < cap ? nchunks : cap);
}

extern "C" int tribe_adaptive_avg_pool_fwd(const float* x, float* y, int64_t rows, int64_t t_in, int64_t t_out, void* stream) {
  if (!x || !y || rows <= 0 || t_in <= 0 || t_out <= 0) return set_error(TRIBE_EINVAL, "pool_fwd: bad arguments");
  if ((reinterpret_cast<uintptr_t>(x) & 15)) return set_error(TRIBE_EINVAL, "pool_fwd: input must be 16-byte aligned");
  const size_t smem = sizeof(float) * (2 * kPoolRows * t_in + 2 * t_out);
  if (smem > 200 * 1024) return set_error(TRIBE_EINVAL, "pool_fwd: t_in too large");
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(pool_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  pool_fwd_kernel<<<pool_grid(rows, smem), 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, rows, static_cast<int>(t_in), static_cast<int>(t_out));
  TRIBE_CHECK_LAUNCH("pool_fwd");
  return TRIBE_OK;
}

extern "C" int tribe_adaptive_avg_pool_bwd(const float* dy, float* dx, int64_t rows, int64_t t_in, int64_t t_out, void* stream) {
  if (!dy || !dx || rows <= 0 || t_in <= 0 || t_out <= 0) return set_error(TRIBE_EINVAL, "pool_bwd: bad arguments");
  if ((reinterpret_cast<uintptr_t>(dy) & 15)) return set_error(TRIBE_EINVAL, "pool_bwd: input must be 16-byte aligned");
  const size_t smem = sizeof(float) * (2 * kPoolRows * t_out + t_out + 2 * t_in);
  if (smem > 200 * 1024) return set_error(TRIBE_EINVAL, "pool_bwd: t_out too large");
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(pool_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr = true;
  }
  pool_bwd_kernel<<<pool_grid(rows, smem), 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(dy, dx, rows, static_cast<int>(t_in), static_cast<int>(t_out));
  TRIBE_CHECK_LAUNCH("pool_bwd");
  return TRIBE_OK;
}
