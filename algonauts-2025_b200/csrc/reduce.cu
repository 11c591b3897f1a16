// Loss / evaluation reductions of the TRIBE hot path (sm_100a): fused MSE forward+gradient and per-parcel Pearson
// sufficient statistics.  HBM-bound (8 B per parcel-TR: one fp32 prediction + one fp32 target, single pass);
// warp-shuffle + shared-memory reductions, fp64 cross-chunk accumulation.
//
// reference: nn.MSELoss at algonauts2025/pl_module.py:56; torchmetrics PearsonCorrCoef updates at pl_module.py:93-106
// (restated in oracle/tm_pearson.py); the final scipy.stats.pearsonr loop at algonauts2025/main.py:459-477.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "tribe_b200.h"
#include "tribe_internal.h"

namespace tribe {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------ MSE
constexpr int kMsePartials = 1024;

__global__ void __launch_bounds__(256) mse_partial_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                          float* __restrict__ grad, float gscale, int64_t n, double* __restrict__ partial) {
  __shared__ double red[8];
  const int64_t nvec = n >> 2;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(target) | reinterpret_cast<uintptr_t>(grad)) & 15) == 0;
  double acc = 0.0;
  if (vec) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
      const float4 p = __ldg(reinterpret_cast<const float4*>(pred) + i);
      const float4 t = __ldg(reinterpret_cast<const float4*>(target) + i);
      const float dx = p.x - t.x, dy = p.y - t.y, dz = p.z - t.z, dw = p.w - t.w;
      acc += static_cast<double>(dx * dx + dy * dy + dz * dz + dw * dw);
      if (grad) reinterpret_cast<float4*>(grad)[i] = make_float4(dx * gscale, dy * gscale, dz * gscale, dw * gscale);
    }
    for (int64_t i = (nvec << 2) + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
      const float d = pred[i] - target[i];
      acc += static_cast<double>(d * d);
      if (grad) grad[i] = d * gscale;
    }
  } else {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
      const float d = pred[i] - target[i];
      acc += static_cast<double>(d * d);
      if (grad) grad[i] = d * gscale;
    }
  }
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < 8 ? red[threadIdx.x] : 0.0;
    v = warp_sum_d(v);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
  }
}

__global__ void __launch_bounds__(256) mse_final_kernel(const double* __restrict__ partial, int nparts, int64_t n, float* __restrict__ loss) {
  __shared__ double red[8];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) acc += partial[i];
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < 8 ? red[threadIdx.x] : 0.0;
    v = warp_sum_d(v);
    if (threadIdx.x == 0) loss[0] = static_cast<float>(v / static_cast<double>(n));
  }
}

// ------------------------------------------------------------------------------------------------ Pearson statistics
// Fast path: row-major (n_rows, n_parcels) fp32 matrices.  Thread = 4 consecutive parcels (one float4 of pred and one of
// target per row), block = one row-chunk x 1024 parcels; fp32 partial sums over kInner rows are folded into fp64
// accumulators, block results are added to stats[6][n_parcels] with fp64 atomics.
constexpr int kInner = 8;

__global__ void __launch_bounds__(256) pearson_rowmajor_kernel(const float* __restrict__ pred, const float* __restrict__ target, int64_t n_rows,
                                                               int64_t n_parcels, int64_t rows_per_block, const long long* __restrict__ group,
                                                               const float* __restrict__ shift, double* __restrict__ stats) {
  const int64_t p0 = (static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x) * 4;
  if (p0 >= n_parcels) return;
  float kx[4] = {0, 0, 0, 0}, ky[4] = {0, 0, 0, 0};  // per-parcel pivots: sums of (x - kx), (y - ky) keep fp32 products well conditioned
  if (shift) {
    for (int j = 0; j < 4; ++j)
      if (p0 + j < n_parcels) kx[j] = shift[p0 + j], ky[j] = shift[n_parcels + p0 + j];
  }
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_block;
  const int64_t r1 = min(n_rows, r0 + rows_per_block);
  if (r0 >= r1) return;
  const bool full4 = p0 + 4 <= n_parcels;
  double sx[4] = {0, 0, 0, 0}, sy[4] = {0, 0, 0, 0}, sxx[4] = {0, 0, 0, 0}, syy[4] = {0, 0, 0, 0}, sxy[4] = {0, 0, 0, 0};
  long long cur_group = group ? group[r0] : 0;
  int64_t n_in_group = 0;
  auto flush = [&](long long gsel, int64_t count) {
    double* st = stats + gsel * 6 * n_parcels;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (p0 + j < n_parcels) {
        atomicAdd(st + 0 * n_parcels + p0 + j, static_cast<double>(count));
        atomicAdd(st + 1 * n_parcels + p0 + j, sx[j]);
        atomicAdd(st + 2 * n_parcels + p0 + j, sy[j]);
        atomicAdd(st + 3 * n_parcels + p0 + j, sxx[j]);
        atomicAdd(st + 4 * n_parcels + p0 + j, syy[j]);
        atomicAdd(st + 5 * n_parcels + p0 + j, sxy[j]);
      }
      sx[j] = sy[j] = sxx[j] = syy[j] = sxy[j] = 0.0;
    }
  };
  for (int64_t r = r0; r < r1; r += kInner) {
    const int nr = static_cast<int>(min(static_cast<int64_t>(kInner), r1 - r));
    if (group) {
      // a chunk of rows never straddles groups in the accumulators: flush when the group id changes
      bool same = true;
      for (int k = 0; k < nr; ++k) same &= (group[r + k] == cur_group);
      if (!same) {
        for (int k = 0; k < nr; ++k) {
          const long long gk = group[r + k];
          if (gk != cur_group) {
            flush(cur_group, n_in_group);
            cur_group = gk, n_in_group = 0;
          }
          float x[4] = {0, 0, 0, 0}, y[4] = {0, 0, 0, 0};
          for (int j = 0; j < 4; ++j)
            if (p0 + j < n_parcels) x[j] = pred[(r + k) * n_parcels + p0 + j] - kx[j], y[j] = target[(r + k) * n_parcels + p0 + j] - ky[j];
          for (int j = 0; j < 4; ++j) sx[j] += x[j], sy[j] += y[j], sxx[j] += (double)x[j] * x[j], syy[j] += (double)y[j] * y[j], sxy[j] += (double)x[j] * y[j];
          ++n_in_group;
        }
        continue;
      }
    }
    float fx[4] = {0, 0, 0, 0}, fy[4] = {0, 0, 0, 0}, fxx[4] = {0, 0, 0, 0}, fyy[4] = {0, 0, 0, 0}, fxy[4] = {0, 0, 0, 0};
    if (full4 && ((n_parcels & 3) == 0)) {
      float4 xs[kInner], ys[kInner];
#pragma unroll
      for (int k = 0; k < kInner; ++k) {
        if (k < nr) {
          xs[k] = __ldg(reinterpret_cast<const float4*>(pred + (r + k) * n_parcels + p0));
          ys[k] = __ldg(reinterpret_cast<const float4*>(target + (r + k) * n_parcels + p0));
        }
      }
#pragma unroll
      for (int k = 0; k < kInner; ++k) {
        if (k < nr) {
          const float x[4] = {xs[k].x - kx[0], xs[k].y - kx[1], xs[k].z - kx[2], xs[k].w - kx[3]};
          const float y[4] = {ys[k].x - ky[0], ys[k].y - ky[1], ys[k].z - ky[2], ys[k].w - ky[3]};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            fx[j] += x[j], fy[j] += y[j];
            fxx[j] = fmaf(x[j], x[j], fxx[j]), fyy[j] = fmaf(y[j], y[j], fyy[j]), fxy[j] = fmaf(x[j], y[j], fxy[j]);
          }
        }
      }
    } else {
      for (int k = 0; k < nr; ++k) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (p0 + j < n_parcels) {
            const float x = pred[(r + k) * n_parcels + p0 + j] - kx[j], y = target[(r + k) * n_parcels + p0 + j] - ky[j];
            fx[j] += x, fy[j] += y, fxx[j] = fmaf(x, x, fxx[j]), fyy[j] = fmaf(y, y, fyy[j]), fxy[j] = fmaf(x, y, fxy[j]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) sx[j] += fx[j], sy[j] += fy[j], sxx[j] += fxx[j], syy[j] += fyy[j], sxy[j] += fxy[j];
    n_in_group += nr;
  }
  flush(cur_group, n_in_group);
}

// Strided path: element(r, p) = base[(r / t_len) * stride_b + p * stride_p + (r % t_len) * stride_t]  (the flattened
// "(b t) d" view of a (B, D, T) prediction tensor, pl_module.py:54-55).  One warp per (b, parcel): lanes run along t.
__global__ void __launch_bounds__(256) pearson_strided_kernel(const float* __restrict__ pred, const float* __restrict__ target, int64_t n_b,
                                                              int64_t n_parcels, int64_t t_len, int64_t stride_b, int64_t stride_p, int64_t stride_t,
                                                              const long long* __restrict__ group, const float* __restrict__ shift,
                                                              double* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const int64_t p = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int64_t b = blockIdx.y;
  if (p >= n_parcels || b >= n_b) return;
  const float kx = shift ? shift[p] : 0.f, ky = shift ? shift[n_parcels + p] : 0.f;
  const float* xp = pred + b * stride_b + p * stride_p;
  const float* yp = target + b * stride_b + p * stride_p;
  float fx = 0, fy = 0, fxx = 0, fyy = 0, fxy = 0;
  double dx = 0, dy = 0, dxx = 0, dyy = 0, dxy = 0;
  int cnt = 0;
  for (int64_t t = lane; t < t_len; t += 32) {
    const float x = __ldg(xp + t * stride_t) - kx, y = __ldg(yp + t * stride_t) - ky;
    fx += x, fy += y, fxx = fmaf(x, x, fxx), fyy = fmaf(y, y, fyy), fxy = fmaf(x, y, fxy);
    if (++cnt == 16) {
      dx += fx, dy += fy, dxx += fxx, dyy += fyy, dxy += fxy;
      fx = fy = fxx = fyy = fxy = 0.f, cnt = 0;
    }
  }
  dx += fx, dy += fy, dxx += fxx, dyy += fyy, dxy += fxy;
  dx = warp_sum_d(dx), dy = warp_sum_d(dy), dxx = warp_sum_d(dxx), dyy = warp_sum_d(dyy), dxy = warp_sum_d(dxy);
  if (lane == 0) {
    double* st = stats + (group ? group[b] : 0) * 6 * n_parcels;
    atomicAdd(st + 0 * n_parcels + p, static_cast<double>(t_len));
    atomicAdd(st + 1 * n_parcels + p, dx);
    atomicAdd(st + 2 * n_parcels + p, dy);
    atomicAdd(st + 3 * n_parcels + p, dxx);
    atomicAdd(st + 4 * n_parcels + p, dyy);
    atomicAdd(st + 5 * n_parcels + p, dxy);
  }
}

// Fast (B, D, T) path: t contiguous and parcels adjacent (stride_t == 1, stride_p == t_len), i.e. the prediction tensor
// exactly as FmriEncoder.forward returns it, or a parcel slice [:, lo:hi] of it (stride_b stays the full D*T).  A block
// owns PB consecutive parcels (PB * t_len / VEC <= 256 items) and a chunk of windows b; thread i keeps the SAME
// (parcel, t) position for every b, so all loads of a block at one b are one contiguous 16 B-per-lane run and the only
// cross-thread reduction is one shared-memory fold per block.  kBdtUnroll windows of independent loads are in flight.
constexpr int kBdtUnroll = 4;

template <int VEC>
__global__ void __launch_bounds__(256) pearson_bdt_kernel(const float* __restrict__ pred, const float* __restrict__ target, int64_t n_b,
                                                          int64_t n_parcels, int t_len, int64_t stride_b, int pb, int64_t b_per_block,
                                                          const long long* __restrict__ group, const float* __restrict__ shift,
                                                          double* __restrict__ stats) {
  __shared__ double sh[64 * 5];  // pb <= 64 parcels x 5 sums
  const int tv = t_len / VEC;
  const int item = threadIdx.x;
  const int pl = item / tv;  // parcel within the block
  const int64_t p = static_cast<int64_t>(blockIdx.x) * pb + pl;
  const bool active = pl < pb && p < n_parcels;
  const int64_t b0 = static_cast<int64_t>(blockIdx.y) * b_per_block;
  const int64_t b1 = min(n_b, b0 + b_per_block);
  if (b0 >= b1) return;
  const int64_t off = static_cast<int64_t>(blockIdx.x) * pb * t_len + static_cast<int64_t>(item) * VEC;
  // per-parcel pivot (a thread keeps ONE parcel for its whole life): the products below are of O(sigma) values even when
  // |mean| >> sigma, which raw fp32 moments cannot represent (scipy / torchmetrics centre before multiplying)
  const float kx = (shift && active) ? shift[p] : 0.f, ky = (shift && active) ? shift[n_parcels + p] : 0.f;
  double sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
  long long cur = group ? group[b0] : 0;
  int64_t cnt = 0;

  auto flush = [&](long long gsel, int64_t count) {  // block-uniform call sites only
    for (int i = threadIdx.x; i < pb * 5; i += blockDim.x) sh[i] = 0.0;
    __syncthreads();
    if (active) {
      atomicAdd(&sh[pl * 5 + 0], sx), atomicAdd(&sh[pl * 5 + 1], sy), atomicAdd(&sh[pl * 5 + 2], sxx);
      atomicAdd(&sh[pl * 5 + 3], syy), atomicAdd(&sh[pl * 5 + 4], sxy);
    }
    __syncthreads();
    double* st = stats + gsel * 6 * n_parcels;
    for (int i = threadIdx.x; i < pb * 6; i += blockDim.x) {
      const int q = i / 6, k = i - q * 6;
      const int64_t pp = static_cast<int64_t>(blockIdx.x) * pb + q;
      if (pp < n_parcels) atomicAdd(st + k * n_parcels + pp, k == 0 ? static_cast<double>(count * t_len) : sh[q * 5 + k - 1]);
    }
    __syncthreads();
    sx = sy = sxx = syy = sxy = 0.0;
  };

  int64_t b = b0;
  while (b < b1) {
    int nk = static_cast<int>(min(static_cast<int64_t>(kBdtUnroll), b1 - b));
    if (group) {
      const long long gid = group[b];
      if (gid != cur) {
        flush(cur, cnt);
        cur = gid, cnt = 0;
      }
      int run = 1;
      while (run < nk && group[b + run] == gid) ++run;
      nk = run;
    }
    if (active) {
      float fx = 0, fy = 0, fxx = 0, fyy = 0, fxy = 0;
      if (VEC == 4) {
        float4 xs[kBdtUnroll], ys[kBdtUnroll];
#pragma unroll
        for (int k = 0; k < kBdtUnroll; ++k) {
          if (k < nk) {
            xs[k] = __ldcs(reinterpret_cast<const float4*>(pred + (b + k) * stride_b + off));
            ys[k] = __ldcs(reinterpret_cast<const float4*>(target + (b + k) * stride_b + off));
          }
        }
#pragma unroll
        for (int k = 0; k < kBdtUnroll; ++k) {
          if (k < nk) {
            const float x[4] = {xs[k].x - kx, xs[k].y - kx, xs[k].z - kx, xs[k].w - kx};
            const float y[4] = {ys[k].x - ky, ys[k].y - ky, ys[k].z - ky, ys[k].w - ky};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              fx += x[j], fy += y[j];
              fxx = fmaf(x[j], x[j], fxx), fyy = fmaf(y[j], y[j], fyy), fxy = fmaf(x[j], y[j], fxy);
            }
          }
        }
      } else {
        float xs[kBdtUnroll], ys[kBdtUnroll];
#pragma unroll
        for (int k = 0; k < kBdtUnroll; ++k) {
          if (k < nk) xs[k] = __ldcs(pred + (b + k) * stride_b + off), ys[k] = __ldcs(target + (b + k) * stride_b + off);
        }
#pragma unroll
        for (int k = 0; k < kBdtUnroll; ++k) {
          if (k < nk) {
            const float x = xs[k] - kx, y = ys[k] - ky;
            fx += x, fy += y, fxx = fmaf(x, x, fxx), fyy = fmaf(y, y, fyy), fxy = fmaf(x, y, fxy);
          }
        }
      }
      sx += fx, sy += fy, sxx += fxx, syy += fyy, sxy += fxy;
    }
    b += nk, cnt += nk;
  }
  flush(cur, cnt);
}

// pivot = the FIRST sample of every parcel (row 0): within a few sigma of the mean for any sane column, and EXACTLY the
// column for a constant one, whose shifted sums are then exactly zero -> r = 0/0 = NaN like scipy's constant-input result
__global__ void __launch_bounds__(256) pearson_pick_shift_kernel(const float* __restrict__ pred, const float* __restrict__ target, int64_t n_parcels,
                                                                 int64_t stride_p, float* __restrict__ shift) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p < n_parcels) shift[p] = pred[p * stride_p], shift[n_parcels + p] = target[p * stride_p];
}

// sums about pivot k -> sums about pivot k' (fp64; d = k - k'):  S1' = S1 + n d,  S2' = S2 + 2 d S1 + n d^2,
// Sxy' = Sxy + dx Sy + dy Sx + n dx dy.  Used to merge statistics taken with different pivots (ranks, checkpoints).
__global__ void __launch_bounds__(256) pearson_recenter_kernel(double* __restrict__ stats, int64_t n_groups, int64_t n_parcels,
                                                               const float* __restrict__ shift_old, const float* __restrict__ shift_new) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_groups * n_parcels) return;
  const int64_t g = i / n_parcels, p = i - g * n_parcels;
  double* st = stats + g * 6 * n_parcels;
  const double dx = (shift_old ? static_cast<double>(shift_old[p]) : 0.0) - (shift_new ? static_cast<double>(shift_new[p]) : 0.0);
  const double dy = (shift_old ? static_cast<double>(shift_old[n_parcels + p]) : 0.0) - (shift_new ? static_cast<double>(shift_new[n_parcels + p]) : 0.0);
  const double n = st[p], sx = st[n_parcels + p], sy = st[2 * n_parcels + p];
  st[3 * n_parcels + p] += 2.0 * dx * sx + n * dx * dx;
  st[4 * n_parcels + p] += 2.0 * dy * sy + n * dy * dy;
  st[5 * n_parcels + p] += dx * sy + dy * sx + n * dx * dy;
  st[n_parcels + p] = sx + n * dx;
  st[2 * n_parcels + p] = sy + n * dy;
}

__global__ void __launch_bounds__(256) pearson_finalize_kernel(const double* __restrict__ stats, int64_t n_parcels, float* __restrict__ r_out,
                                                               float* __restrict__ mean_out) {
  __shared__ double red[8];
  double acc = 0.0;
  for (int64_t p = threadIdx.x; p < n_parcels; p += blockDim.x) {
    const double n = stats[p], sx = stats[n_parcels + p], sy = stats[2 * n_parcels + p];
    const double sxx = stats[3 * n_parcels + p], syy = stats[4 * n_parcels + p], sxy = stats[5 * n_parcels + p];
    const double cov = sxy - sx * sy / n, vx = sxx - sx * sx / n, vy = syy - sy * sy / n;
    double r = cov / sqrt(vx * vy);
    // NaN (constant column: var == 0 -> 0/0; NaN input) propagates like scipy / torchmetrics.  fmin/fmax would
    // silently turn it into -1 (they return the non-NaN operand), so clamp only ordered values.
    if (r > 1.0) r = 1.0;
    if (r < -1.0) r = -1.0;
    if (r_out) r_out[p] = static_cast<float>(r);
    acc += r;
  }
  acc = warp_sum_d(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < 8 ? red[threadIdx.x] : 0.0;
    v = warp_sum_d(v);
    if (threadIdx.x == 0 && mean_out) mean_out[0] = static_cast<float>(v / static_cast<double>(n_parcels));
  }
}

}  // namespace tribe

using namespace tribe;

extern "C" int tribe_mse_fwd_bwd(const float* pred, const float* target, float* loss_out, float* grad, float grad_scale, int64_t n,
                                 double* partial, void* stream) {
  if (!pred || !target || !loss_out || !partial || n <= 0) return set_error(TRIBE_EINVAL, "mse: bad arguments");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int grid = grid_for(n / 4 + 1, 256 * 4, kMsePartials);
  const float gscale = grad_scale * 2.0f / static_cast<float>(n);
  mse_partial_kernel<<<grid, 256, 0, s>>>(pred, target, grad, gscale, n, partial);
  TRIBE_CHECK_LAUNCH("mse_partial");
  mse_final_kernel<<<1, 256, 0, s>>>(partial, grid, n, loss_out);
  TRIBE_CHECK_LAUNCH("mse_final");
  return TRIBE_OK;
}

extern "C" int tribe_pearson_stats(const float* pred, const float* target, int64_t n_rows, int64_t n_parcels, int64_t t_len, int64_t stride_b,
                                   int64_t stride_p, int64_t stride_t, const int64_t* group, int64_t n_groups, const float* shift, double* stats,
                                   void* stream) {
  if (!pred || !target || !stats || n_rows <= 0 || n_parcels <= 0 || t_len <= 0 || n_rows % t_len)
    return set_error(TRIBE_EINVAL, "pearson_stats: bad arguments (n_rows must be a multiple of t_len)");
  if (group && n_groups <= 0) return set_error(TRIBE_EINVAL, "pearson_stats: group given without n_groups");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long* g = reinterpret_cast<const long long*>(group);
  if (t_len == 1 && stride_p == 1 && stride_b == n_parcels) {
    const int64_t pblocks = (n_parcels + 1023) / 1024;
    int64_t chunks = (148 * 8 + pblocks - 1) / pblocks;
    const int64_t max_chunks = (n_rows + 4 * kInner - 1) / (4 * kInner);
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks > 65535) chunks = 65535;
    int64_t rpb = (n_rows + chunks - 1) / chunks;
    rpb = (rpb + kInner - 1) / kInner * kInner;
    dim3 grid(static_cast<unsigned>(pblocks), static_cast<unsigned>((n_rows + rpb - 1) / rpb));
    pearson_rowmajor_kernel<<<grid, 256, 0, s>>>(pred, target, n_rows, n_parcels, rpb, g, shift, stats);
    TRIBE_CHECK_LAUNCH("pearson_rowmajor");
  } else if (stride_t == 1 && stride_p == t_len && t_len <= 1024 &&
             (((t_len & 3) == 0 && (stride_b & 3) == 0 && ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(target)) & 15) == 0) ||
              t_len <= 256)) {
    const int64_t n_b = n_rows / t_len;
    const bool vec = (t_len & 3) == 0 && (stride_b & 3) == 0 && ((reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(target)) & 15) == 0;
    const int tv = static_cast<int>(vec ? t_len / 4 : t_len);
    int pb = 256 / tv;
    if (pb > 64) pb = 64;
    const int64_t pblocks = (n_parcels + pb - 1) / pb;
    int64_t chunks = (148 * 8 + pblocks - 1) / pblocks;
    // every chunk of a parcel block ends in fp64 atomics on the SAME 6 x pb addresses: narrow parcel shards (1000 / 8
    // parcels per GPU) must not buy their parallelism with ~100-way atomic contention
    static const int64_t chunk_cap = [] {
      const char* e = getenv("TRIBE_PEARSON_MAX_CHUNKS");
      const int v = e ? atoi(e) : 0;
      return static_cast<int64_t>(v > 0 ? v : 24);
    }();
    if (chunks > chunk_cap) chunks = chunk_cap;
    const int64_t max_chunks = (n_b + kBdtUnroll - 1) / kBdtUnroll;
    if (chunks > max_chunks) chunks = max_chunks;
    if (chunks > 65535) chunks = 65535;
    int64_t bpb = (n_b + chunks - 1) / chunks;
    bpb = (bpb + kBdtUnroll - 1) / kBdtUnroll * kBdtUnroll;
    dim3 grid(static_cast<unsigned>(pblocks), static_cast<unsigned>((n_b + bpb - 1) / bpb));
    if (vec)
      pearson_bdt_kernel<4><<<grid, 256, 0, s>>>(pred, target, n_b, n_parcels, static_cast<int>(t_len), stride_b, pb, bpb, g, shift, stats);
    else
      pearson_bdt_kernel<1><<<grid, 256, 0, s>>>(pred, target, n_b, n_parcels, static_cast<int>(t_len), stride_b, pb, bpb, g, shift, stats);
    TRIBE_CHECK_LAUNCH("pearson_bdt");
  } else {
    const int64_t n_b = n_rows / t_len;
    if (n_b > 65535) return set_error(TRIBE_EINVAL, "pearson_stats: more than 65535 strided blocks per call");
    dim3 grid(static_cast<unsigned>((n_parcels + 7) / 8), static_cast<unsigned>(n_b));
    pearson_strided_kernel<<<grid, 256, 0, s>>>(pred, target, n_b, n_parcels, t_len, stride_b, stride_p, stride_t, g, shift, stats);
    TRIBE_CHECK_LAUNCH("pearson_strided");
  }
  return TRIBE_OK;
}

extern "C" int tribe_pearson_pick_shift(const float* pred, const float* target, int64_t n_parcels, int64_t stride_p, float* shift, void* stream) {
  if (!pred || !target || !shift || n_parcels <= 0 || stride_p <= 0) return set_error(TRIBE_EINVAL, "pearson_pick_shift: bad arguments");
  pearson_pick_shift_kernel<<<static_cast<unsigned>((n_parcels + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pred, target, n_parcels,
                                                                                                                             stride_p, shift);
  TRIBE_CHECK_LAUNCH("pearson_pick_shift");
  return TRIBE_OK;
}

extern "C" int tribe_pearson_recenter(double* stats, int64_t n_groups, int64_t n_parcels, const float* shift_old, const float* shift_new, void* stream) {
  if (!stats || n_groups <= 0 || n_parcels <= 0) return set_error(TRIBE_EINVAL, "pearson_recenter: bad arguments");
  if (!shift_old && !shift_new) return TRIBE_OK;
  const int64_t n = n_groups * n_parcels;
  pearson_recenter_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(stats, n_groups, n_parcels, shift_old,
                                                                                                                   shift_new);
  TRIBE_CHECK_LAUNCH("pearson_recenter");
  return TRIBE_OK;
}

extern "C" int tribe_pearson_finalize(const double* stats, int64_t n_parcels, float* r, float* mean_out, void* stream) {
  if (!stats || n_parcels <= 0 || (!r && !mean_out)) return set_error(TRIBE_EINVAL, "pearson_finalize: bad arguments");
  pearson_finalize_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(stats, n_parcels, r, mean_out);
  TRIBE_CHECK_LAUNCH("pearson_finalize");
  return TRIBE_OK;
}
