// Symmetric InfoNCE over flattened (B*T) rows (reference: algonauts2025/model.py:208-221, used by the contrastive
// branch pl_module.py:59-77).  The (n x n) logits q^ k^T / tau come from the tcgen05 GEMM; these kernels do the two
// cross-entropies and their gradient in single passes over the logits.  Because q^ and k^ are unit vectors,
// |logit| <= 1/tau, so exp(l - 1/tau) cannot overflow and a constant shift replaces the per-row/column max.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tribe_b200.h"
#include "tribe_internal.h"

namespace tribe {

constexpr int kNceMaxPerThread = 32;  // columns per thread at 256 threads -> n <= 8192

__device__ __forceinline__ float nce_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// row_sum[i] = sum_j exp(l_ij - shift);  col_sum[j] += sum_i exp(l_ij - shift)   (one read of the logits)
__global__ void __launch_bounds__(256) nce_expsums_kernel(const float* __restrict__ logits, int n, int64_t ld, float shift,
                                                          float* __restrict__ row_sum, float* __restrict__ col_sum) {
  __shared__ float red[8];
  float cacc[kNceMaxPerThread];
#pragma unroll
  for (int k = 0; k < kNceMaxPerThread; ++k) cacc[k] = 0.f;
  for (int r = blockIdx.x; r < n; r += gridDim.x) {
    const float* lr = logits + static_cast<int64_t>(r) * ld;
    float rs = 0.f;
#pragma unroll
    for (int k = 0; k < kNceMaxPerThread; ++k) {
      const int c = threadIdx.x + k * 256;
      if (c < n) {
        const float e = __expf(__ldg(lr + c) - shift);
        rs += e;
        cacc[k] += e;
      }
    }
    rs = nce_warp_sum(rs);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = rs;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w];
      row_sum[r] = t;
    }
  }
#pragma unroll
  for (int k = 0; k < kNceMaxPerThread; ++k) {
    const int c = threadIdx.x + k * 256;
    if (c < n) atomicAdd(col_sum + c, cacc[k]);
  }
}

// loss = 0.5/n * sum_i [(shift + log row_sum_i) - l_ii] + 0.5/n * sum_j [(shift + log col_sum_j) - l_jj]
__global__ void __launch_bounds__(256) nce_loss_kernel(const float* __restrict__ logits, int n, int64_t ld, float shift,
                                                       const float* __restrict__ row_sum, const float* __restrict__ col_sum,
                                                       float* __restrict__ loss) {
  __shared__ double red[8];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = logits[static_cast<int64_t>(i) * ld + i];
    acc += (static_cast<double>(shift) + log(static_cast<double>(row_sum[i])) - d) + (static_cast<double>(shift) + log(static_cast<double>(col_sum[i])) - d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    loss[0] = static_cast<float>(0.5 * t / static_cast<double>(n));
  }
}

// G_ij = scale * (e_ij / row_sum_i + e_ij / col_sum_j - 2 delta_ij), bf16, padded columns zeroed.
__global__ void __launch_bounds__(256) nce_grad_kernel(const float* __restrict__ logits, int n, int64_t ld, float shift,
                                                       const float* __restrict__ row_sum, const float* __restrict__ col_sum,
                                                       const float* __restrict__ upstream, float scale, __nv_bfloat16* __restrict__ g, int64_t ldg) {
  const float s = scale * (upstream ? __ldg(upstream) : 1.0f);
  for (int r = blockIdx.x; r < n; r += gridDim.x) {
    const float* lr = logits + static_cast<int64_t>(r) * ld;
    __nv_bfloat16* gr = g + static_cast<int64_t>(r) * ldg;
    const float inv_rs = 1.0f / __ldg(row_sum + r);
    for (int c = threadIdx.x; c < ldg; c += blockDim.x) {
      float v = 0.f;
      if (c < n) {
        const float e = __expf(__ldg(lr + c) - shift);
        v = s * (e * inv_rs + e / __ldg(col_sum + c) - (c == r ? 2.0f : 0.0f));
      }
      gr[c] = __float2bfloat16(v);
    }
  }
}

__global__ void __launch_bounds__(256) cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int64_t n) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t nvec = n >> 3;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src) + i);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
    const float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]), c = __bfloat1622float2(h[2]), d = __bfloat1622float2(h[3]);
    reinterpret_cast<float4*>(dst)[2 * i] = make_float4(a.x, a.y, b.x, b.y);
    reinterpret_cast<float4*>(dst)[2 * i + 1] = make_float4(c.x, c.y, d.x, d.y);
  }
  for (int64_t i = (nvec << 3) + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = __bfloat162float(src[i]);
}

}  // namespace tribe

using namespace tribe;

extern "C" int tribe_nce_expsums(const float* logits, int64_t n, int64_t ld, float shift, float* row_sum, float* col_sum, void* stream) {
  if (!logits || !row_sum || !col_sum || n <= 0 || ld < n || n > 256 * kNceMaxPerThread) return set_error(TRIBE_EINVAL, "nce_expsums: bad arguments (n <= 8192)");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(col_sum, 0, sizeof(float) * n, s);
  if (e != cudaSuccess) return set_cuda_error(e, "nce_expsums memset");
  nce_expsums_kernel<<<grid_for(n, 8, 148 * 4), 256, 0, s>>>(logits, static_cast<int>(n), ld, shift, row_sum, col_sum);
  TRIBE_CHECK_LAUNCH("nce_expsums");
  return TRIBE_OK;
}

extern "C" int tribe_nce_loss(const float* logits, int64_t n, int64_t ld, float shift, const float* row_sum, const float* col_sum, float* loss_out,
                              void* stream) {
  if (!logits || !row_sum || !col_sum || !loss_out || n <= 0 || ld < n) return set_error(TRIBE_EINVAL, "nce_loss: bad arguments");
  nce_loss_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(logits, static_cast<int>(n), ld, shift, row_sum, col_sum, loss_out);
  TRIBE_CHECK_LAUNCH("nce_loss");
  return TRIBE_OK;
}

extern "C" int tribe_nce_grad(const float* logits, int64_t n, int64_t ld, float shift, const float* row_sum, const float* col_sum,
                              const float* upstream, float scale, void* g_bf16, int64_t ldg, void* stream) {
  if (!logits || !row_sum || !col_sum || !g_bf16 || n <= 0 || ld < n || ldg < n) return set_error(TRIBE_EINVAL, "nce_grad: bad arguments");
  nce_grad_kernel<<<grid_for(n, 1, 148 * 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      logits, static_cast<int>(n), ld, shift, row_sum, col_sum, upstream, scale, reinterpret_cast<__nv_bfloat16*>(g_bf16), ldg);
  TRIBE_CHECK_LAUNCH("nce_grad");
  return TRIBE_OK;
}

extern "C" int tribe_cast_bf16_f32(const void* src_bf16, float* dst, int64_t n, void* stream) {
  if (!src_bf16 || !dst || n <= 0) return set_error(TRIBE_EINVAL, "cast_bf16_f32: bad arguments");
  if ((reinterpret_cast<uintptr_t>(src_bf16) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15)) return set_error(TRIBE_EINVAL, "cast_bf16_f32: 16-byte alignment required");
  cast_bf16_f32_kernel<<<grid_for(n / 8 + 1, 256, 148 * 16), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(src_bf16), dst, n);
  TRIBE_CHECK_LAUNCH("cast_bf16_f32");
  return TRIBE_OK;
}
