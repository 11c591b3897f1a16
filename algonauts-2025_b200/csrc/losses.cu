// Alternative training losses of the TRIBE grids (algonauts2025/grids/run_ensemble.py:29 samples
// MSELoss | PearsonLoss | SmoothL1Loss | HuberLoss; built by modeling_utils/losses/base.py:43-59) on the same reduction
// structure as the fused MSE: one pass that produces the loss value AND the gradient w.r.t. the prediction.
//
//  * point-wise losses (torch.nn.SmoothL1Loss / HuberLoss / L1Loss, reduction="mean"): 8 B read + 4 B written per element;
//  * PearsonLoss (modeling_utils/losses/losses.py:11-42, dim=1: one correlation per parcel over all (b, t) rows):
//    forward = the Pearson sufficient-statistics kernel (reduce.cu, 8 B / parcel-TR) + a per-parcel finalize that emits
//    the loss and 4 coefficients per parcel; backward = one element-wise pass (8 B read + 4 B written).
#include <cuda_runtime.h>
#include <stdint.h>

#include "tribe_b200.h"
#include "tribe_internal.h"

namespace tribe {

__device__ __forceinline__ double loss_warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// inv = 1 / prm (0 when prm == 0), hoisted out of the element loop: two fp32 divisions per element made the kernel
// issue-bound (66 % of HBM peak) instead of bandwidth-bound
template <int KIND>
__device__ __forceinline__ void point_loss(float d, float prm, float inv, float& l, float& g) {
  const float a = fabsf(d);
  if (KIND == TRIBE_LOSS_SMOOTH_L1) {  // torch.nn.SmoothL1Loss(beta=prm); beta == 0 degenerates to L1
    if (a < prm) {
      g = d * inv, l = 0.5f * d * g;
    } else {
      l = a - 0.5f * prm, g = (d > 0.f) - (d < 0.f);
    }
  } else if (KIND == TRIBE_LOSS_HUBER) {  // torch.nn.HuberLoss(delta=prm)
    if (a <= prm) {
      l = 0.5f * d * d, g = d;
    } else {
      l = prm * (a - 0.5f * prm), g = prm * ((d > 0.f) - (d < 0.f));
    }
  } else {  // torch.nn.L1Loss
    l = a, g = (d > 0.f) - (d < 0.f);
  }
}

constexpr int kLossPartials = 1024;

template <int KIND>
__global__ void __launch_bounds__(256) point_loss_partial_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                                 float* __restrict__ grad, float gscale, float prm, int64_t n,
                                                                 double* __restrict__ partial) {
  __shared__ double red[8];
  const int64_t nvec = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(target) | reinterpret_cast<uintptr_t>(grad)) & 15) == 0;
  double acc = 0.0;
  const float inv = prm > 0.f ? 1.0f / prm : 0.f;
  const int64_t first = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (vec) {
    for (int64_t i = first; i < nvec; i += stride) {
      const float4 p = __ldcs(reinterpret_cast<const float4*>(pred) + i);
      const float4 t = __ldcs(reinterpret_cast<const float4*>(target) + i);
      float l0, l1, l2, l3, g0, g1, g2, g3;
      point_loss<KIND>(p.x - t.x, prm, inv, l0, g0), point_loss<KIND>(p.y - t.y, prm, inv, l1, g1);
      point_loss<KIND>(p.z - t.z, prm, inv, l2, g2), point_loss<KIND>(p.w - t.w, prm, inv, l3, g3);
      acc += static_cast<double>((l0 + l1) + (l2 + l3));
      if (grad) reinterpret_cast<float4*>(grad)[i] = make_float4(g0 * gscale, g1 * gscale, g2 * gscale, g3 * gscale);
    }
  }
  for (int64_t i = (vec ? (nvec << 2) : 0) + first; i < n; i += stride) {
    float l, g;
    point_loss<KIND>(pred[i] - target[i], prm, inv, l, g);
    acc += static_cast<double>(l);
    if (grad) grad[i] = g * gscale;
  }
  acc = loss_warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < 8 ? red[threadIdx.x] : 0.0;
    v = loss_warp_sum(v);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
  }
}

__global__ void __launch_bounds__(256) loss_final_kernel(const double* __restrict__ partial, int nparts, double inv_n, float* __restrict__ loss) {
  __shared__ double red[8];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) acc += partial[i];
  acc = loss_warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < 8 ? red[threadIdx.x] : 0.0;
    v = loss_warp_sum(v);
    if (threadIdx.x == 0) loss[0] = static_cast<float>(v * inv_n);
  }
}

// PearsonLoss finalize: per parcel, from stats = [n, Sx, Sy, Sxx, Syy, Sxy] (fp64):
//   a = Sxy - Sx Sy / n,  sx = sqrt(Sxx - Sx^2/n),  sy likewise,  D = sx sy + 1e-8,  pcc = a / D,  loss_p = 1 - pcc
//   d pcc / d x_i = (y_i - my) / D - a sy (x_i - mx) / (D^2 sx)        (centering is its own adjoint here)
// coef[0..3][p] = mx, my, c1 = 1/D, c2 = a sy / (D^2 sx)   (c2 = 0 for a constant prediction, like autograd's 0 * inf guard
// does NOT do — the reference would produce NaN there; we keep NaN out of the weights and document it).
__global__ void __launch_bounds__(256) pearson_loss_finalize_kernel(const double* __restrict__ stats, int64_t n_parcels, const float* __restrict__ shift,
                                                                    float* __restrict__ coef, float* __restrict__ loss, double scale) {
  __shared__ double red[8];
  double acc = 0.0;
  for (int64_t p = threadIdx.x; p < n_parcels; p += blockDim.x) {
    const double n = stats[p], sx1 = stats[n_parcels + p], sy1 = stats[2 * n_parcels + p];
    const double sxx = stats[3 * n_parcels + p], syy = stats[4 * n_parcels + p], sxy = stats[5 * n_parcels + p];
    const double a = sxy - sx1 * sy1 / n;
    const double sx = sqrt(fmax(sxx - sx1 * sx1 / n, 0.0)), sy = sqrt(fmax(syy - sy1 * sy1 / n, 0.0));
    const double D = sx * sy + 1e-8;
    acc += 1.0 - a / D;
    if (coef) {
      // the statistics are sums of (x - pivot): the means the backward centres with are pivot + S1 / n
      coef[p] = static_cast<float>((shift ? static_cast<double>(shift[p]) : 0.0) + sx1 / n);
      coef[n_parcels + p] = static_cast<float>((shift ? static_cast<double>(shift[n_parcels + p]) : 0.0) + sy1 / n);
      coef[2 * n_parcels + p] = static_cast<float>(1.0 / D);
      coef[3 * n_parcels + p] = sx > 0.0 ? static_cast<float>(a * sy / (D * D * sx)) : 0.f;
    }
  }
  acc = loss_warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < 8 ? red[threadIdx.x] : 0.0;
    v = loss_warp_sum(v);
    if (threadIdx.x == 0) loss[0] = static_cast<float>(v * scale);
  }
}

// grad[i] = -(upstream * scale) * (c1[p] (y_i - my[p]) - c2[p] (x_i - mx[p])),  p = (i / t_len) % n_parcels
// (t_len = 1: row-major (N, O); t_len = T: contiguous (B, O, T) — the (b t) d rearrange is never materialised).
__global__ void __launch_bounds__(256) pearson_loss_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                               const float* __restrict__ coef, const float* __restrict__ upstream, float scale,
                                                               float* __restrict__ grad, int64_t n, int64_t n_parcels, int64_t t_len) {
  const float up = -scale * (upstream ? upstream[0] : 1.f);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const bool vec = (t_len % 4 == 0 || (t_len == 1 && n_parcels % 4 == 0)) &&
                   ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(target) | reinterpret_cast<uintptr_t>(grad)) & 15) == 0;
  if (vec && t_len > 1) {  // four consecutive t of one parcel
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < (n >> 2); i += stride) {
      const int64_t p = ((i << 2) / t_len) % n_parcels;
      const float mx = __ldg(coef + p), my = __ldg(coef + n_parcels + p), c1 = __ldg(coef + 2 * n_parcels + p) * up,
                  c2 = __ldg(coef + 3 * n_parcels + p) * up;
      const float4 x = __ldcs(reinterpret_cast<const float4*>(pred) + i), y = __ldcs(reinterpret_cast<const float4*>(target) + i);
      __stcs(reinterpret_cast<float4*>(grad) + i, make_float4(c1 * (y.x - my) - c2 * (x.x - mx), c1 * (y.y - my) - c2 * (x.y - mx),
                                                             c1 * (y.z - my) - c2 * (x.z - mx), c1 * (y.w - my) - c2 * (x.w - mx)));
    }
  } else if (vec) {  // row-major: four consecutive parcels
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < (n >> 2); i += stride) {
      const int64_t p = (i << 2) % n_parcels;
      const float4 mx = __ldg(reinterpret_cast<const float4*>(coef + p)), my = __ldg(reinterpret_cast<const float4*>(coef + n_parcels + p));
      const float4 c1 = __ldg(reinterpret_cast<const float4*>(coef + 2 * n_parcels + p)), c2 = __ldg(reinterpret_cast<const float4*>(coef + 3 * n_parcels + p));
      const float4 x = __ldcs(reinterpret_cast<const float4*>(pred) + i), y = __ldcs(reinterpret_cast<const float4*>(target) + i);
      __stcs(reinterpret_cast<float4*>(grad) + i,
             make_float4(up * (c1.x * (y.x - my.x) - c2.x * (x.x - mx.x)), up * (c1.y * (y.y - my.y) - c2.y * (x.y - mx.y)),
                         up * (c1.z * (y.z - my.z) - c2.z * (x.z - mx.z)), up * (c1.w * (y.w - my.w) - c2.w * (x.w - mx.w))));
    }
  } else {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
      const int64_t p = (i / t_len) % n_parcels;
      grad[i] = up * (coef[2 * n_parcels + p] * (target[i] - coef[n_parcels + p]) - coef[3 * n_parcels + p] * (pred[i] - coef[p]));
    }
  }
}

// dst[i] = src[i] * scalar[0]  (autograd's upstream gradient of a fused loss is a DEVICE scalar: no host read)
__global__ void __launch_bounds__(256) scale_dev_kernel(const float* __restrict__ src, const float* __restrict__ scalar, float* __restrict__ dst,
                                                        int64_t n) {
  const float a = __ldg(scalar);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  const int64_t nvec = vec ? (n >> 2) : 0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const float4 x = __ldcs(reinterpret_cast<const float4*>(src) + i);
    __stcs(reinterpret_cast<float4*>(dst) + i, make_float4(x.x * a, x.y * a, x.z * a, x.w * a));
  }
  for (int64_t i = (nvec << 2) + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i] * a;
}

}  // namespace tribe

using namespace tribe;

extern "C" int tribe_point_loss_fwd_bwd(const float* pred, const float* target, float* loss_out, float* grad, int32_t kind, float param,
                                        float grad_scale, int64_t n, double* partial, void* stream) {
  if (!pred || !target || !loss_out || !partial || n <= 0) return set_error(TRIBE_EINVAL, "point_loss: bad arguments");
  if (param < 0.f) return set_error(TRIBE_EINVAL, "point_loss: beta / delta must be >= 0");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = grid_for(n / 4 + 1, 256 * 4, kLossPartials);
  const float gscale = grad_scale / static_cast<float>(n);
  switch (kind) {
    case TRIBE_LOSS_SMOOTH_L1: point_loss_partial_kernel<TRIBE_LOSS_SMOOTH_L1><<<grid, 256, 0, s>>>(pred, target, grad, gscale, param, n, partial); break;
    case TRIBE_LOSS_HUBER: point_loss_partial_kernel<TRIBE_LOSS_HUBER><<<grid, 256, 0, s>>>(pred, target, grad, gscale, param, n, partial); break;
    case TRIBE_LOSS_L1: point_loss_partial_kernel<TRIBE_LOSS_L1><<<grid, 256, 0, s>>>(pred, target, grad, gscale, param, n, partial); break;
    default: return set_error(TRIBE_EINVAL, "point_loss: unknown kind");
  }
  TRIBE_CHECK_LAUNCH("point_loss_partial");
  loss_final_kernel<<<1, 256, 0, s>>>(partial, grid, 1.0 / static_cast<double>(n), loss_out);
  TRIBE_CHECK_LAUNCH("loss_final");
  return TRIBE_OK;
}

extern "C" int tribe_pearson_loss_finalize(const double* stats, int64_t n_parcels, int32_t reduction_mean, const float* shift, float* coef,
                                           float* loss_out, void* stream) {
  if (!stats || !loss_out || n_parcels <= 0) return set_error(TRIBE_EINVAL, "pearson_loss_finalize: bad arguments");
  pearson_loss_finalize_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(stats, n_parcels, shift, coef, loss_out,
                                                                                     reduction_mean ? 1.0 / static_cast<double>(n_parcels) : 1.0);
  TRIBE_CHECK_LAUNCH("pearson_loss_finalize");
  return TRIBE_OK;
}

extern "C" int tribe_pearson_loss_bwd(const float* pred, const float* target, const float* coef, const float* upstream, int32_t reduction_mean,
                                      float* grad, int64_t n, int64_t n_parcels, int64_t t_len, void* stream) {
  if (!pred || !target || !coef || !grad || n <= 0 || n_parcels <= 0 || t_len <= 0 || n % (n_parcels * t_len))
    return set_error(TRIBE_EINVAL, "pearson_loss_bwd: bad arguments (n must be a multiple of n_parcels * t_len)");
  const float scale = reduction_mean ? 1.0f / static_cast<float>(n_parcels) : 1.0f;
  pearson_loss_bwd_kernel<<<grid_for(n / 4 + 1, 256, 148 * 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pred, target, coef, upstream, scale,
                                                                                                               grad, n, n_parcels, t_len);
  TRIBE_CHECK_LAUNCH("pearson_loss_bwd");
  return TRIBE_OK;
}

extern "C" int tribe_scale_dev(const float* src, const float* scalar_dev, float* dst, int64_t n, void* stream) {
  if (!src || !scalar_dev || !dst || n <= 0) return set_error(TRIBE_EINVAL, "scale_dev: bad arguments");
  scale_dev_kernel<<<grid_for(n / 4 + 1, 256, 148 * 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, scalar_dev, dst, n);
  TRIBE_CHECK_LAUNCH("scale_dev");
  return TRIBE_OK;
}

extern "C" int tribe_memset_zero(void* ptr, int64_t n_bytes, void* stream) {
  if (!ptr || n_bytes < 0) return set_error(TRIBE_EINVAL, "memset_zero: bad arguments");
  if (n_bytes == 0) return TRIBE_OK;
  cudaError_t e = cudaMemsetAsync(ptr, 0, static_cast<size_t>(n_bytes), reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return set_cuda_error(e, "memset_zero");
  return TRIBE_OK;
}
