// Fused Adam step over a flat fp32 parameter range that ALSO emits the bf16 shadow weights the tensor cores read
// (one pass: 16 B read + 14 B written per parameter instead of torch's multi-tensor Adam followed by a separate cast).
// Same arithmetic as torch.optim.Adam (amsgrad=False, maximize=False), reference recipe
// algonauts2025/grids/defaults.py:126-141 (Adam lr 1e-4, weight_decay 0, OneCycleLR stepping per batch).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "adam_math.cuh"
#include "tribe_b200.h"
#include "tribe_internal.h"

namespace tribe {

__device__ __forceinline__ uint32_t opt_pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// hyper == nullptr: scalars come from the launch arguments; otherwise from 6 floats in device memory
// {beta1, beta2, step_size, inv_bc2_sqrt, eps, weight_decay}, so a captured CUDA graph replays with fresh values.
// All traffic is streaming (evict-first): nothing here is re-read before ~28 GB have passed through L2, and the
// kernel may run beside the backward GEMMs whose operand tiles should stay L2-resident.
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                   __nv_bfloat16* __restrict__ p16, int64_t n, float beta1, float beta2, float step_size,
                                                   float inv_bc2_sqrt, float eps, float wd, const float* __restrict__ hyper) {
  TRIBE_PDL_ENTRY();
  if (hyper) beta1 = hyper[0], beta2 = hyper[1], step_size = hyper[2], inv_bc2_sqrt = hyper[3], eps = hyper[4], wd = hyper[5];
  const int64_t nvec = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float4 pp = __ldcs(reinterpret_cast<const float4*>(p) + i);
    const float4 gg = __ldcs(reinterpret_cast<const float4*>(g) + i);
    float4 mm = __ldcs(reinterpret_cast<const float4*>(m) + i);
    float4 vv = __ldcs(reinterpret_cast<const float4*>(v) + i);
    adam_one(pp.x, gg.x, mm.x, vv.x, beta1, beta2, step_size, inv_bc2_sqrt, eps, wd);
    adam_one(pp.y, gg.y, mm.y, vv.y, beta1, beta2, step_size, inv_bc2_sqrt, eps, wd);
    adam_one(pp.z, gg.z, mm.z, vv.z, beta1, beta2, step_size, inv_bc2_sqrt, eps, wd);
    adam_one(pp.w, gg.w, mm.w, vv.w, beta1, beta2, step_size, inv_bc2_sqrt, eps, wd);
    __stcs(reinterpret_cast<float4*>(p) + i, pp);
    __stcs(reinterpret_cast<float4*>(m) + i, mm);
    __stcs(reinterpret_cast<float4*>(v) + i, vv);
    if (p16) __stcs(reinterpret_cast<uint2*>(p16) + i, make_uint2(opt_pack2(pp.x, pp.y), opt_pack2(pp.z, pp.w)));
  }
  for (int64_t i = (nvec << 2) + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    float pp = p[i], mm = m[i], vv = v[i];
    adam_one(pp, g[i], mm, vv, beta1, beta2, step_size, inv_bc2_sqrt, eps, wd);
    p[i] = pp, m[i] = mm, v[i] = vv;
    if (p16) p16[i] = __float2bfloat16(pp);
  }
}

__global__ void set_floats_kernel(float* __restrict__ dst, int n, float v0, float v1, float v2, float v3, float v4, float v5, float v6, float v7) {
  const float v[8] = {v0, v1, v2, v3, v4, v5, v6, v7};
  if (threadIdx.x < n) dst[threadIdx.x] = v[threadIdx.x];
}

constexpr int kHyperBatch = 96;
struct HyperBatch {
  int32_t slot[kHyperBatch];
  float val[kHyperBatch][6];
};
// one launch refreshes up to 96 hyper-parameter blocks (block i of `base` = 8 floats at base + 8 * slot[i])
__global__ void set_hyper_batch_kernel(float* __restrict__ base, const __grid_constant__ HyperBatch hb, int n) {
  const int i = threadIdx.x / 8, j = threadIdx.x % 8;
  if (i < n && j < 6) base[static_cast<int64_t>(hb.slot[i]) * 8 + j] = hb.val[i][j];
}

}  // namespace tribe

extern "C" int tribe_adam_step(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, double lr, double beta1, double beta2,
                               double eps, double weight_decay, int64_t step, int32_t max_blocks, void* stream) {
  using namespace tribe;
  if (!p || !g || !m || !v || n <= 0 || step <= 0) return set_error(TRIBE_EINVAL, "adam_step: bad arguments");
  const uintptr_t al = reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v);
  if ((al & 15) || (reinterpret_cast<uintptr_t>(p_bf16) & 7)) return set_error(TRIBE_EINVAL, "adam_step: buffers must be 16-byte aligned (bf16: 8)");
  // hyper-parameters arrive as doubles (Python floats) and the bias corrections are formed in fp64 like torch does
  // (torch/optim/adam.py: bias_correction = 1 - beta ** step with Python floats): 1 - 0.999^k is ill-conditioned for small k
  const double bc1 = 1.0 - pow(beta1, static_cast<double>(step));
  const double bc2 = 1.0 - pow(beta2, static_cast<double>(step));
  const float step_size = static_cast<float>(lr / bc1);
  const float inv_bc2_sqrt = static_cast<float>(1.0 / sqrt(bc2));
  launch_k(adam_kernel, dim3(grid_for(n / 4 + 1, 256, max_blocks > 0 ? max_blocks : 148 * 8)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), p, g,
           m, v, reinterpret_cast<__nv_bfloat16*>(p_bf16), n, static_cast<float>(beta1), static_cast<float>(beta2), step_size, inv_bc2_sqrt,
           static_cast<float>(eps), static_cast<float>(weight_decay), static_cast<const float*>(nullptr));
  TRIBE_CHECK_LAUNCH("adam_step");
  return TRIBE_OK;
}

extern "C" int tribe_adam_step_dev(float* p, const float* g, float* m, float* v, void* p_bf16, int64_t n, const float* hyper, int32_t max_blocks, void* stream) {
  using namespace tribe;
  if (!p || !g || !m || !v || !hyper || n <= 0) return set_error(TRIBE_EINVAL, "adam_step_dev: bad arguments");
  const uintptr_t al = reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v);
  if ((al & 15) || (reinterpret_cast<uintptr_t>(p_bf16) & 7)) return set_error(TRIBE_EINVAL, "adam_step_dev: buffers must be 16-byte aligned (bf16: 8)");
  launch_k(adam_kernel, dim3(grid_for(n / 4 + 1, 256, max_blocks > 0 ? max_blocks : 148 * 8)), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream), p, g,
           m, v, reinterpret_cast<__nv_bfloat16*>(p_bf16), n, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, hyper);
  TRIBE_CHECK_LAUNCH("adam_step_dev");
  return TRIBE_OK;
}

extern "C" int tribe_adam_hyper(float* hyper_dev, double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step,
                                void* stream) {
  using namespace tribe;
  if (!hyper_dev || step <= 0) return set_error(TRIBE_EINVAL, "adam_hyper: bad arguments");
  const double bc1 = 1.0 - pow(beta1, static_cast<double>(step));
  const double bc2 = 1.0 - pow(beta2, static_cast<double>(step));
  set_floats_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(hyper_dev, 6, static_cast<float>(beta1), static_cast<float>(beta2),
                                                                        static_cast<float>(lr / bc1), static_cast<float>(1.0 / sqrt(bc2)),
                                                                        static_cast<float>(eps), static_cast<float>(weight_decay), 0.f, 0.f);
  TRIBE_CHECK_LAUNCH("adam_hyper");
  return TRIBE_OK;
}

extern "C" int tribe_adam_hyper_batch(float* hyper_base, const int32_t* slots_host, const int64_t* steps_host, const double* lr_host,
                                      const double* beta1_host, const double* beta2_host, const double* eps_host, const double* wd_host, int32_t n,
                                      void* stream) {
  using namespace tribe;
  if (!hyper_base || n < 0 || (n > 0 && (!slots_host || !steps_host || !lr_host || !beta1_host || !beta2_host || !eps_host || !wd_host)))
    return set_error(TRIBE_EINVAL, "adam_hyper_batch: bad arguments");
  for (int32_t base = 0; base < n; base += kHyperBatch) {
    const int cnt = n - base < kHyperBatch ? n - base : kHyperBatch;
    HyperBatch hb;
    for (int i = 0; i < cnt; ++i) {
      const int q = base + i;
      if (steps_host[q] <= 0 || slots_host[q] < 0) return set_error(TRIBE_EINVAL, "adam_hyper_batch: step must be >= 1 and slot >= 0");
      const double bc1 = 1.0 - pow(beta1_host[q], static_cast<double>(steps_host[q]));
      const double bc2 = 1.0 - pow(beta2_host[q], static_cast<double>(steps_host[q]));
      hb.slot[i] = slots_host[q];
      hb.val[i][0] = static_cast<float>(beta1_host[q]), hb.val[i][1] = static_cast<float>(beta2_host[q]);
      hb.val[i][2] = static_cast<float>(lr_host[q] / bc1), hb.val[i][3] = static_cast<float>(1.0 / sqrt(bc2));
      hb.val[i][4] = static_cast<float>(eps_host[q]), hb.val[i][5] = static_cast<float>(wd_host[q]);
    }
    set_hyper_batch_kernel<<<1, kHyperBatch * 8, 0, reinterpret_cast<cudaStream_t>(stream)>>>(hyper_base, hb, cnt);
    TRIBE_CHECK_LAUNCH("adam_hyper_batch");
  }
  return TRIBE_OK;
}
