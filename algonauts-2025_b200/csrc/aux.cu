// The steps either side of the hot path, all HBM-bound byte movers / small reductions:
//  * window assembly   data_utils/base.py:167-198 (TimedArray overlap slicing) + data_utils/segments.py:144-180: gather the
//                      [start, start+T) slice of every window out of device-resident timeline feature arrays, zero padded;
//  * ensemble average  algonauts2025/grids/average_submissions.py:107-125: per-parcel softmax(r / tau) weights over the
//                      members, weighted mean of their predictions;
//  * retrieval ranks   modeling_utils/metrics/metrics.py:66-121 (Rank._compute_sim norm_kind="y", _compute_ranks) behind
//                      TopkAcc (metrics.py:194-218), used as val/retrieval_top1 (grids/defaults.py:119-123);
//  * SWA               running average of the flat parameter buffer (algonauts2025/main.py:365-373, torch AveragedModel:
//                      avg += (p - avg) / (n_averaged + 1)).
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "tribe_b200.h"
#include "tribe_internal.h"

namespace tribe {

// ------------------------------------------------------------------------------------------------ window gather
// out[b, r, t] = (dst0[b] <= t < dst0[b] + len[b]) ? src[b][r * t_total[b] + src0[b] + (t - dst0[b])] : 0
// for r in [0, rows) (rows = L*D feature rows, or parcels for the fMRI target), t in [0, T).  One block = (chunk of t,
// 8 rows, window b); lanes run along t so both the strided source reads and the destination writes are contiguous runs.
template <typename SRC>
__global__ void __launch_bounds__(256) gather_windows_kernel(const SRC* const* __restrict__ src, const long long* __restrict__ t_total,
                                                             const int* __restrict__ dst0, const int* __restrict__ src0, const int* __restrict__ len,
                                                             float* __restrict__ out, int64_t rows, int T) {
  const int b = blockIdx.z;
  const int64_t r = static_cast<int64_t>(blockIdx.y) * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const SRC* s = src[b] + r * t_total[b];
  float* o = out + (static_cast<int64_t>(b) * rows + r) * T;
  const int d0 = dst0[b], s0 = src0[b], n = len[b];
  const int lane = threadIdx.x & 31;
  // one warp per row; 8 independent (unaligned, hence scalar) loads in flight per lane before the first store
  for (int base = 0; base < T; base += 256) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = base + j * 32 + lane - d0;
      v[j] = (k >= 0 && k < n && base + j * 32 + lane < T) ? static_cast<float>(__ldcs(s + s0 + k)) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int t = base + j * 32 + lane;
      if (t < T) __stcs(o + t, v[j]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ ensemble average
// axis 1: w[m, :] = softmax over the PARCELS of member m — what the reference computes (`pearsons.softmax(dim=1)` on the
// (n_submissions, n_voxels) matrix, average_submissions.py:108-109).  One block per member.
__global__ void __launch_bounds__(256) ensemble_weights_rows_kernel(const float* __restrict__ r, int64_t O, float inv_tau, float* __restrict__ w) {
  __shared__ float red[8];
  __shared__ float bc;
  const float* rm = r + static_cast<int64_t>(blockIdx.x) * O;
  float* wm = w + static_cast<int64_t>(blockIdx.x) * O;
  float mx = -CUDART_INF_F;
  for (int64_t p = threadIdx.x; p < O; p += blockDim.x) mx = fmaxf(mx, rm[p] * inv_tau);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = red[0];
    for (int i = 1; i < 8; ++i) v = fmaxf(v, red[i]);
    bc = v;
  }
  __syncthreads();
  mx = bc;
  float sum = 0.f;
  for (int64_t p = threadIdx.x; p < O; p += blockDim.x) sum += expf(rm[p] * inv_tau - mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int i = 0; i < 8; ++i) v += red[i];
    bc = v;
  }
  __syncthreads();
  const float inv = 1.0f / bc;
  for (int64_t p = threadIdx.x; p < O; p += blockDim.x) wm[p] = expf(rm[p] * inv_tau - mx) * inv;
}

// axis 0: w[:, p] = softmax over the MEMBERS for parcel p (weights of one parcel sum to one).
__global__ void __launch_bounds__(256) ensemble_weights_kernel(const float* __restrict__ r, int M, int64_t O, float inv_tau, float* __restrict__ w) {
  const int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= O) return;
  float mx = -CUDART_INF_F;
  for (int m = 0; m < M; ++m) mx = fmaxf(mx, r[m * O + p] * inv_tau);
  float sum = 0.f;
  for (int m = 0; m < M; ++m) sum += expf(r[m * O + p] * inv_tau - mx);
  for (int m = 0; m < M; ++m) w[m * O + p] = expf(r[m * O + p] * inv_tau - mx) / sum;
}

// out[n, p] = sum_m w[m, p] * preds[m, n, p]   (preds stacked (M, N, O) row-major; w == nullptr: plain mean).
// Thread = 4 consecutive parcels, 4 row lanes per block; the block's weight tile lives in shared memory.
constexpr int kEnsMaxMembers = 48;
__global__ void __launch_bounds__(256) ensemble_average_kernel(const float* __restrict__ preds, const float* __restrict__ w, int M, int64_t N,
                                                               int64_t O, int64_t rows_per_block, float* __restrict__ out) {
  __shared__ float ws[kEnsMaxMembers][256];
  const int cg = threadIdx.x & 63, lane_r = threadIdx.x >> 6;
  const int64_t c0 = static_cast<int64_t>(blockIdx.x) * 256;
  for (int i = threadIdx.x; i < M * 256; i += 256) {
    const int m = i >> 8, c = i & 255;
    ws[m][c] = (c0 + c < O) ? (w ? w[m * O + c0 + c] : 1.0f / static_cast<float>(M)) : 0.f;
  }
  __syncthreads();
  const int64_t c = c0 + cg * 4;
  if (c >= O) return;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_block, r1 = min(N, r0 + rows_per_block);
  const bool vec = (O & 3) == 0 && ((reinterpret_cast<uintptr_t>(preds) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  for (int64_t r = r0 + lane_r; r < r1; r += 4) {
    if (vec) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int m = 0; m < M; ++m) {
        const float4 x = __ldcs(reinterpret_cast<const float4*>(preds + (static_cast<int64_t>(m) * N + r) * O + c));
        acc.x = fmaf(ws[m][cg * 4 + 0], x.x, acc.x), acc.y = fmaf(ws[m][cg * 4 + 1], x.y, acc.y);
        acc.z = fmaf(ws[m][cg * 4 + 2], x.z, acc.z), acc.w = fmaf(ws[m][cg * 4 + 3], x.w, acc.w);
      }
      __stcs(reinterpret_cast<float4*>(out + r * O + c), acc);
    } else {
      for (int j = 0; j < 4 && c + j < O; ++j) {
        float acc = 0.f;
        for (int m = 0; m < M; ++m) acc = fmaf(ws[m][cg * 4 + j], preds[(static_cast<int64_t>(m) * N + r) * O + c + j], acc);
        out[r * O + c + j] = acc;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ retrieval ranks
// y[row] = mean over the last dim (the `.mean(dim=-1)` of pl_module.py:100-101 on (B, O, T) tensors); warp per row.
__global__ void __launch_bounds__(256) mean_lastdim_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t rows, int t) {
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float acc = 0.f;
  for (int i = threadIdx.x & 31; i < t; i += 32) acc += x[row * t + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) y[row] = acc / static_cast<float>(t);
}

// scores[b, o] = <x_b, y_o> / (1e-15 + ||y_o||)                                  (Rank._compute_sim, norm_kind="y")
// ranks[b] = (#{o: s[b,o] > s[b,b]} + #{o: s[b,o] >= s[b,b]} - 1) / 2, comparisons with NaN count as false (nansum);
// a negative result (NaN true score) becomes n // 2                            (Rank._compute_ranks without labels).
// One block per query b; warps stride over the candidates o, each computing one dot product + norm at a time.
__global__ void __launch_bounds__(256) retrieval_ranks_kernel(const float* __restrict__ x, const float* __restrict__ y, int n, int c,
                                                              float* __restrict__ ranks, float* __restrict__ scores_out) {
  extern __shared__ float sc[];  // n scores of this query
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* xb = x + static_cast<int64_t>(b) * c;
  for (int o = warp; o < n; o += 8) {
    const float* yo = y + static_cast<int64_t>(o) * c;
    float dot = 0.f, nn = 0.f;
    for (int i = lane; i < c; i += 32) {
      const float yv = yo[i];
      dot = fmaf(xb[i], yv, dot), nn = fmaf(yv, yv, nn);
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, k), nn += __shfl_xor_sync(0xffffffffu, nn, k);
    if (lane == 0) {
      sc[o] = dot * (1.0f / (1e-15f + sqrtf(nn)));
      if (scores_out) scores_out[static_cast<int64_t>(b) * n + o] = sc[o];
    }
  }
  __syncthreads();
  __shared__ int cnt[2];
  if (threadIdx.x < 2) cnt[threadIdx.x] = 0;
  __syncthreads();
  const float truth = sc[b];
  int gt = 0, ge = 0;
  for (int o = threadIdx.x; o < n; o += blockDim.x) gt += sc[o] > truth, ge += sc[o] >= truth;
  atomicAdd(&cnt[0], gt), atomicAdd(&cnt[1], ge);
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = 0.5f * static_cast<float>(cnt[0] + cnt[1] - 1);
    if (r < 0.f) r = static_cast<float>(n / 2);
    ranks[b] = r;
  }
}

// ------------------------------------------------------------------------------------------------ (B, R, C) -> (B, C, R)
// The materialised `rearrange("b d t -> (b t) d")` of pl_module.py:54-55 (needed only by losses / metrics that are not
// fused) and the per-window `pred.T` of the submission assembly (callbacks.py:66): 32 x 32 tiles through padded smem,
// coalesced on both sides, bit-exact.
__global__ void __launch_bounds__(256) transpose_last2_kernel(const float* __restrict__ x, float* __restrict__ y, int R, int C) {
  __shared__ float tile[32][33];
  const int64_t base = static_cast<int64_t>(blockIdx.z) * R * C;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int r = r0 + ty + j, c = c0 + tx;
    if (r < R && c < C) tile[ty + j][tx] = x[base + static_cast<int64_t>(r) * C + c];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const int c = c0 + ty + j, r = r0 + tx;
    if (r < R && c < C) y[base + static_cast<int64_t>(c) * R + r] = tile[tx][ty + j];
  }
}

// ------------------------------------------------------------------------------------------------ SWA
__global__ void __launch_bounds__(256) swa_update_kernel(float* __restrict__ avg, const float* __restrict__ p, float inv_n1, int64_t n) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t nvec = n >> 2;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float4 a = __ldcs(reinterpret_cast<const float4*>(avg) + i);
    const float4 x = __ldcs(reinterpret_cast<const float4*>(p) + i);
    a.x += (x.x - a.x) * inv_n1, a.y += (x.y - a.y) * inv_n1, a.z += (x.z - a.z) * inv_n1, a.w += (x.w - a.w) * inv_n1;
    __stcs(reinterpret_cast<float4*>(avg) + i, a);
  }
  for (int64_t i = (nvec << 2) + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    avg[i] += (p[i] - avg[i]) * inv_n1;
}

}  // namespace tribe

using namespace tribe;

extern "C" int tribe_gather_windows(const void* const* src_ptrs, int32_t src_dtype, const int64_t* t_total, const int32_t* dst_start,
                                    const int32_t* src_start, const int32_t* length, float* out, int64_t n_windows, int64_t rows, int64_t t_out,
                                    void* stream) {
  if (!src_ptrs || !t_total || !dst_start || !src_start || !length || !out || n_windows <= 0 || rows <= 0 || t_out <= 0)
    return set_error(TRIBE_EINVAL, "gather_windows: bad arguments");
  if (n_windows > 65535 || (rows + 7) / 8 > 65535) return set_error(TRIBE_EINVAL, "gather_windows: too many windows / rows for one launch");
  dim3 grid(1, static_cast<unsigned>((rows + 7) / 8), static_cast<unsigned>(n_windows));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const long long* tt = reinterpret_cast<const long long*>(t_total);
  if (src_dtype == TRIBE_DT_F32)
    gather_windows_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const float* const*>(src_ptrs), tt, dst_start, src_start, length, out, rows,
                                                      static_cast<int>(t_out));
  else if (src_dtype == TRIBE_DT_F64)
    gather_windows_kernel<double><<<grid, 256, 0, s>>>(reinterpret_cast<const double* const*>(src_ptrs), tt, dst_start, src_start, length, out, rows,
                                                       static_cast<int>(t_out));
  else
    return set_error(TRIBE_EINVAL, "gather_windows: source dtype must be f32 or f64");
  TRIBE_CHECK_LAUNCH("gather_windows");
  return TRIBE_OK;
}

extern "C" int tribe_ensemble_weights(const float* r, int64_t n_members, int64_t n_parcels, float temperature, int32_t axis, float* w,
                                      void* stream) {
  if (!r || !w || n_members <= 0 || n_parcels <= 0 || !(temperature > 0.f) || (axis != 0 && axis != 1))
    return set_error(TRIBE_EINVAL, "ensemble_weights: bad arguments");
  if (axis == 1)
    ensemble_weights_rows_kernel<<<static_cast<unsigned>(n_members), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(r, n_parcels,
                                                                                                                     1.0f / temperature, w);
  else
    ensemble_weights_kernel<<<static_cast<unsigned>((n_parcels + 255) / 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        r, static_cast<int>(n_members), n_parcels, 1.0f / temperature, w);
  TRIBE_CHECK_LAUNCH("ensemble_weights");
  return TRIBE_OK;
}

extern "C" int tribe_ensemble_average(const float* preds, const float* w, int64_t n_members, int64_t n_rows, int64_t n_parcels, float* out,
                                      void* stream) {
  if (!preds || !out || n_members <= 0 || n_rows <= 0 || n_parcels <= 0) return set_error(TRIBE_EINVAL, "ensemble_average: bad arguments");
  if (n_members > kEnsMaxMembers) return set_error(TRIBE_EINVAL, "ensemble_average: more than 48 members per launch");
  const int64_t cblocks = (n_parcels + 255) / 256;
  int64_t chunks = (148 * 8 + cblocks - 1) / cblocks;
  if (chunks > (n_rows + 3) / 4) chunks = (n_rows + 3) / 4;
  if (chunks > 65535) chunks = 65535;
  const int64_t rpb = (n_rows + chunks - 1) / chunks;
  dim3 grid(static_cast<unsigned>(cblocks), static_cast<unsigned>((n_rows + rpb - 1) / rpb));
  ensemble_average_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(preds, w, static_cast<int>(n_members), n_rows, n_parcels, rpb,
                                                                                   out);
  TRIBE_CHECK_LAUNCH("ensemble_average");
  return TRIBE_OK;
}

extern "C" int tribe_mean_lastdim(const float* x, float* y, int64_t rows, int64_t t, void* stream) {
  if (!x || !y || rows <= 0 || t <= 0 || t > INT32_MAX) return set_error(TRIBE_EINVAL, "mean_lastdim: bad arguments");
  mean_lastdim_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, rows, static_cast<int>(t));
  TRIBE_CHECK_LAUNCH("mean_lastdim");
  return TRIBE_OK;
}

extern "C" int tribe_retrieval_ranks(const float* x, const float* y, int64_t n, int64_t c, float* ranks, float* scores_out, void* stream) {
  if (!x || !y || !ranks || n <= 0 || c <= 0) return set_error(TRIBE_EINVAL, "retrieval_ranks: bad arguments");
  if (n > 10000) return set_error(TRIBE_EINVAL, "retrieval_ranks: more than 10000 candidates per call");
  retrieval_ranks_kernel<<<static_cast<unsigned>(n), 256, sizeof(float) * n, reinterpret_cast<cudaStream_t>(stream)>>>(
      x, y, static_cast<int>(n), static_cast<int>(c), ranks, scores_out);
  TRIBE_CHECK_LAUNCH("retrieval_ranks");
  return TRIBE_OK;
}

extern "C" int tribe_swa_update(float* avg, const float* params, int64_t n, int64_t n_averaged, void* stream) {
  if (!avg || !params || n <= 0 || n_averaged < 0) return set_error(TRIBE_EINVAL, "swa_update: bad arguments");
  if ((reinterpret_cast<uintptr_t>(avg) | reinterpret_cast<uintptr_t>(params)) & 15) return set_error(TRIBE_EINVAL, "swa_update: buffers must be 16-byte aligned");
  swa_update_kernel<<<grid_for(n / 4 + 1, 256, 148 * 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      avg, params, 1.0f / static_cast<float>(n_averaged + 1), n);
  TRIBE_CHECK_LAUNCH("swa_update");
  return TRIBE_OK;
}

extern "C" int tribe_transpose_last2(const float* x, float* y, int64_t batch, int64_t rows, int64_t cols, void* stream) {
  if (!x || !y || batch <= 0 || rows <= 0 || cols <= 0 || batch > 65535 || rows > INT32_MAX || cols > INT32_MAX)
    return set_error(TRIBE_EINVAL, "transpose_last2: bad arguments");
  if ((rows + 31) / 32 > 65535) return set_error(TRIBE_EINVAL, "transpose_last2: too many rows for one launch");
  dim3 grid(static_cast<unsigned>((cols + 31) / 32), static_cast<unsigned>((rows + 31) / 32), static_cast<unsigned>(batch));
  transpose_last2_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, static_cast<int>(rows), static_cast<int>(cols));
  TRIBE_CHECK_LAUNCH("transpose_last2");
  return TRIBE_OK;
}
