// HBM-bound kernels of the TRIBE hot path (sm_100a): feature ingest, ScaleNorm forward / sub-layer backward tail,
// attention softmax forward / backward, column sums, casts, adaptive average pooling, small layout converters.
// Each is a coalesced, 128-bit vectorised streaming kernel; roofline = bytes moved / measured HBM copy bandwidth.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "tribe_b200.h"
#include "tribe_internal.h"

namespace tribe {

constexpr int kMaxBlocks = 148 * 16;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Block-wide sum; `red` must hold >= 32 floats.  All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = lane < nw ? red[lane] : 0.0f;
  return warp_sum(t);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
}

// ------------------------------------------------------------------------------------------------ feature ingest
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<double>(double v) { return static_cast<float>(v); }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

// x (B, L, D, T) -> out bf16 (B*T, ld_out): tile of 64 d x 64 t transposed through shared memory.
// reference: algonauts2025/model.py:147-155 (cast, rearrange "b l d t -> b (l d) t", transpose(1, 2)).
template <typename T>
__global__ void __launch_bounds__(256) ingest_kernel(const T* __restrict__ x, int64_t L, int64_t D, int64_t Tn, int layer_mean,
                                                     __nv_bfloat16* __restrict__ out, int64_t ld_out, int64_t col_off) {
  __shared__ float tile[64][65];
  const int b = blockIdx.z;
  const int64_t n_out_rows = layer_mean ? D : L * D;  // output feature index range
  const int64_t f0 = static_cast<int64_t>(blockIdx.y) * 64;
  const int64_t t0 = static_cast<int64_t>(blockIdx.x) * 64;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < 64; r += 8) {
    const int64_t f = f0 + r;
    float v0 = 0.f, v1 = 0.f;
    if (f < n_out_rows) {
      if (!layer_mean) {
        const T* src = x + (static_cast<int64_t>(b) * L * D + f) * Tn;
        if (t0 + lane < Tn) v0 = to_f32<T>(src[t0 + lane]);
        if (t0 + 32 + lane < Tn) v1 = to_f32<T>(src[t0 + 32 + lane]);
      } else {
        for (int64_t l = 0; l < L; ++l) {
          const T* src = x + ((static_cast<int64_t>(b) * L + l) * D + f) * Tn;
          if (t0 + lane < Tn) v0 += to_f32<T>(src[t0 + lane]);
          if (t0 + 32 + lane < Tn) v1 += to_f32<T>(src[t0 + 32 + lane]);
        }
        const float inv = 1.0f / static_cast<float>(L);
        v0 *= inv, v1 *= inv;
      }
    }
    tile[r][lane] = v0;
    tile[r][lane + 32] = v1;
  }
  __syncthreads();
  for (int tt = warp; tt < 64; tt += 8) {
    const int64_t t = t0 + tt;
    if (t >= Tn) break;
    const int64_t f = f0 + 2 * lane;
    __nv_bfloat16* dst = out + (static_cast<int64_t>(b) * Tn + t) * ld_out + col_off + f;
    const float a = tile[2 * lane][tt], c = tile[2 * lane + 1][tt];
    if (f + 1 < n_out_rows && ((reinterpret_cast<uintptr_t>(dst) & 3) == 0)) {
      *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(a, c);
    } else {
      if (f < n_out_rows) dst[0] = __float2bfloat16(a);
      if (f + 1 < n_out_rows) dst[1] = __float2bfloat16(c);
    }
  }
}

// ------------------------------------------------------------------------------------------------ ScaleNorm
// y = x / max(||x||, 1e-12) * sqrt(dim) * g   (oracle/xt_encoder.py ScaleNorm; F.normalize eps)
__global__ void __launch_bounds__(256) scalenorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                            __nv_bfloat16* __restrict__ y, float* __restrict__ rnorm, int64_t rows,
                                                            int dim, float mult, float eps) {
  __shared__ float red[32];
  TRIBE_PDL_ENTRY();
  const int nvec = dim >> 2;
  const float scale_g = mult * __ldg(g);
  for (int64_t row = blockIdx.x; row < rows; row += gridDim.x) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * dim);
    float4 cache[4];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int v = threadIdx.x + i * 256;
      if (v < nvec) {
        cache[i] = __ldg(xr + v);
        ss += cache[i].x * cache[i].x + cache[i].y * cache[i].y + cache[i].z * cache[i].z + cache[i].w * cache[i].w;
      }
    }
    for (int v = threadIdx.x + 1024; v < nvec; v += 256) {
      const float4 c = __ldg(xr + v);
      ss += c.x * c.x + c.y * c.y + c.z * c.z + c.w * c.w;
    }
    ss = block_sum(ss, red);
    const float rn = 1.0f / fmaxf(sqrtf(ss), eps);
    if (threadIdx.x == 0 && rnorm) rnorm[row] = rn;
    const float s = rn * scale_g;
    uint2* yr = reinterpret_cast<uint2*>(y + row * dim);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int v = threadIdx.x + i * 256;
      if (v < nvec) yr[v] = make_uint2(pack_bf16x2(cache[i].x * s, cache[i].y * s), pack_bf16x2(cache[i].z * s, cache[i].w * s));
    }
    for (int v = threadIdx.x + 1024; v < nvec; v += 256) {
      const float4 c = __ldg(xr + v);
      yr[v] = make_uint2(pack_bf16x2(c.x * s, c.y * s), pack_bf16x2(c.z * s, c.w * s));
    }
  }
}

// Warp-per-row variant (dim = 128 * NV): the whole row lives in one warp's registers, every global load of the row is
// issued before the first use and the reduction is five shuffles — no block barrier between the load and store phases,
// eight independent rows in flight per block.
template <int NV>
__global__ void __launch_bounds__(256) scalenorm_fwd_warp_kernel(const float* __restrict__ x, const float* __restrict__ g,
                                                                 __nv_bfloat16* __restrict__ y, float* __restrict__ rnorm, int64_t rows,
                                                                 float mult, float eps) {
  constexpr int dim = NV * 128;
  const int lane = threadIdx.x & 31;
  TRIBE_PDL_ENTRY();
  const float scale_g = mult * __ldg(g);
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5); row < rows; row += static_cast<int64_t>(gridDim.x) * 8) {
    const float4* xr = reinterpret_cast<const float4*>(x + row * dim);
    float4 c[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) c[i] = __ldg(xr + lane + i * 32);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) ss += c[i].x * c[i].x + c[i].y * c[i].y + c[i].z * c[i].z + c[i].w * c[i].w;
    ss = warp_sum(ss);
    const float rn = 1.0f / fmaxf(sqrtf(ss), eps);
    if (lane == 0 && rnorm) rnorm[row] = rn;
    const float sc = rn * scale_g;
    uint2* yr = reinterpret_cast<uint2*>(y + row * dim);
#pragma unroll
    for (int i = 0; i < NV; ++i) yr[lane + i * 32] = make_uint2(pack_bf16x2(c[i].x * sc, c[i].y * sc), pack_bf16x2(c[i].z * sc, c[i].w * sc));
  }
}

// Backward tail of one pre-norm residual sub-layer  x_out = branch(ScaleNorm(x_in)) + x_in * rs :
//   dot      = sum_c d_xn[c] * x_in[c] * rnorm
//   dx_in[c] = sqrt(dim) g rnorm (d_xn[c] - x_in[c] rnorm dot) + dy_out[c] * rs[c]
//   d_rs[c] += dy_out[c] * x_in[c]        d_g += sqrt(dim) * dot
// Each block walks a strided set of rows and keeps its d_rs column partials in registers (dim <= 4096).
// Software-pipelined: while row A is reduced / finished, the loads of the block's next row B are already in flight
// (two register sets, the row loop is unrolled by two), so every block has a row's worth of reads (30 KB at dim 3072)
// outstanding at all times instead of only between a row's first load and its block reduction.
template <int NV>
struct SubRow {
  float4 xc[NV], dyv[NV];
  uint2 dnraw[NV];
};

template <int NV>
__device__ __forceinline__ void sub_row_load(SubRow<NV>& r, const float* __restrict__ dy_out, const __nv_bfloat16* __restrict__ d_xn,
                                             const float* __restrict__ x_in, int64_t row, int dim, int nvec) {
  const float4* xr = reinterpret_cast<const float4*>(x_in + row * dim);
  const float4* dyr = dy_out ? reinterpret_cast<const float4*>(dy_out + row * dim) : nullptr;
  const uint2* dnr = d_xn ? reinterpret_cast<const uint2*>(d_xn + row * dim) : nullptr;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = threadIdx.x + i * 256;
    if (v < nvec) {
      r.xc[i] = __ldcs(xr + v);
      if (dnr) r.dnraw[i] = __ldcs(dnr + v);
      if (dyr) r.dyv[i] = __ldcs(dyr + v);
    }
  }
}

template <int NV>
__device__ __forceinline__ void sub_row_finish(const SubRow<NV>& r, bool has_dn, bool has_dy, const float* __restrict__ rnorm,
                                               const float* __restrict__ rs, float* __restrict__ dx_in, __nv_bfloat16* __restrict__ dx_in_bf16,
                                               int64_t row, int dim, int nvec, float sqrt_dim, float gval, float4 (&rs_acc)[NV], float& dg_acc,
                                               float* red) {
  float4 dn[NV];
  float dot = 0.f;
  if (has_dn) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = threadIdx.x + i * 256;
      if (v < nvec) {
        const float2 a = unpack_bf16x2(r.dnraw[i].x), b = unpack_bf16x2(r.dnraw[i].y);
        dn[i] = make_float4(a.x, a.y, b.x, b.y);
        dot += dn[i].x * r.xc[i].x + dn[i].y * r.xc[i].y + dn[i].z * r.xc[i].z + dn[i].w * r.xc[i].w;
      }
    }
  }
  float coef = 0.f, rn = 0.f;
  if (has_dn) {
    rn = __ldg(rnorm + row);
    dot = block_sum(dot, red) * rn;
    coef = sqrt_dim * gval * rn;
    dg_acc += dot;  // identical in every thread; thread 0 publishes
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = threadIdx.x + i * 256;
    if (v < nvec) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (has_dn) {
        const float k = rn * dot;
        o.x = coef * (dn[i].x - r.xc[i].x * k), o.y = coef * (dn[i].y - r.xc[i].y * k);
        o.z = coef * (dn[i].z - r.xc[i].z * k), o.w = coef * (dn[i].w - r.xc[i].w * k);
      }
      if (has_dy) {
        const float4 dy = r.dyv[i];
        float4 r4 = make_float4(1.f, 1.f, 1.f, 1.f);
        if (rs) r4 = __ldg(reinterpret_cast<const float4*>(rs) + v);
        o.x += dy.x * r4.x, o.y += dy.y * r4.y, o.z += dy.z * r4.z, o.w += dy.w * r4.w;
        rs_acc[i].x += dy.x * r.xc[i].x, rs_acc[i].y += dy.y * r.xc[i].y, rs_acc[i].z += dy.z * r.xc[i].z, rs_acc[i].w += dy.w * r.xc[i].w;
      }
      if (dx_in) __stcs(reinterpret_cast<float4*>(dx_in + row * dim) + v, o);
      // the bf16 copy is the A operand of the next dgrad / wgrad GEMM: keep it cacheable
      if (dx_in_bf16) reinterpret_cast<uint2*>(dx_in_bf16 + row * dim)[v] = make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
    }
  }
}

template <int NV>
__global__ void __launch_bounds__(256, 2) sublayer_bwd_kernel(const float* __restrict__ dy_out, const __nv_bfloat16* __restrict__ d_xn,
                                                              const float* __restrict__ x_in, const float* __restrict__ rnorm,
                                                              const float* __restrict__ g, const float* __restrict__ rs,
                                                              float* __restrict__ dx_in, __nv_bfloat16* __restrict__ dx_in_bf16,
                                                              float* __restrict__ d_rs, float* __restrict__ d_g, int64_t rows, int dim,
                                                              float sqrt_dim /* the norm's constant gain factor: sqrt(dim) or 1 */) {
  __shared__ float red[32];
  TRIBE_PDL_ENTRY();
  const int nvec = dim >> 2;
  const float gval = g ? __ldg(g) : 0.f;
  const bool has_dn = d_xn != nullptr, has_dy = dy_out != nullptr;
  float4 rs_acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) rs_acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  float dg_acc = 0.f;
  SubRow<NV> A, B;
  int64_t row_a = blockIdx.x;
  if (row_a < rows) sub_row_load<NV>(A, dy_out, d_xn, x_in, row_a, dim, nvec);
  while (row_a < rows) {
    const int64_t row_b = row_a + gridDim.x;
    if (row_b < rows) sub_row_load<NV>(B, dy_out, d_xn, x_in, row_b, dim, nvec);
    sub_row_finish<NV>(A, has_dn, has_dy, rnorm, rs, dx_in, dx_in_bf16, row_a, dim, nvec, sqrt_dim, gval, rs_acc, dg_acc, red);
    row_a = row_b + gridDim.x;
    if (row_a < rows) sub_row_load<NV>(A, dy_out, d_xn, x_in, row_a, dim, nvec);
    if (row_b < rows) sub_row_finish<NV>(B, has_dn, has_dy, rnorm, rs, dx_in, dx_in_bf16, row_b, dim, nvec, sqrt_dim, gval, rs_acc, dg_acc, red);
  }
  if (d_rs) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int v = threadIdx.x + i * 256;
      if (v < nvec) {
        atomicAdd(d_rs + 4 * v, rs_acc[i].x), atomicAdd(d_rs + 4 * v + 1, rs_acc[i].y);
        atomicAdd(d_rs + 4 * v + 2, rs_acc[i].z), atomicAdd(d_rs + 4 * v + 3, rs_acc[i].w);
      }
    }
  }
  if (d_g && threadIdx.x == 0) atomicAdd(d_g, sqrt_dim * dg_acc);
}

// ------------------------------------------------------------------------------------------------ half-split rotary
// x_transformers 1.27.x rotary embedding (oracle/xt_encoder.py, semantics "v1.27"): inside every head the rotary dims are
// paired (i, i + rot/2), both rotated by angle pos * inv_freq[i]:  x_i' = x_i cos - x_{i+rot/2} sin * sign,
// x_{i+rot/2}' = x_{i+rot/2} cos + x_i sin * sign   (sign = -1: the transpose, for the backward).  In place on a bf16
// (rows, ld) buffer; a thread owns 8 adjacent i of one (row, head): two 16-byte loads + two 16-byte stores.
// (The interleaved pairing of >= 2.x is fused into the GEMM epilogue — TRIBE_EPI_ROPE; partners 96 columns apart do not
// share an epilogue chunk, nor always a 256-wide tile.)
__global__ void __launch_bounds__(256) rope_half_kernel(__nv_bfloat16* __restrict__ x, int64_t rows, int64_t ld, int64_t col_off, int n_heads,
                                                        int head_dim, int rot_dim, const float2* __restrict__ table, int T, float sign) {
  const int half = rot_dim >> 1, groups = half >> 3;
  const int64_t total = rows * n_heads * groups;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int gq = static_cast<int>(idx % groups);
    const int64_t rest = idx / groups;
    const int h = static_cast<int>(rest % n_heads);
    const int64_t row = rest / n_heads;
    __nv_bfloat16* lo = x + row * ld + col_off + static_cast<int64_t>(h) * head_dim + gq * 8;
    __nv_bfloat16* hi = lo + half;
    const float2* tab = table + static_cast<int64_t>(row % T) * half + gq * 8;
    uint4 a = *reinterpret_cast<const uint4*>(lo), b = *reinterpret_cast<const uint4*>(hi);
    uint32_t* aw = reinterpret_cast<uint32_t*>(&a);
    uint32_t* bw = reinterpret_cast<uint32_t*>(&b);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 x1 = unpack_bf16x2(aw[j]), x2 = unpack_bf16x2(bw[j]);
      const float4 cs = __ldg(reinterpret_cast<const float4*>(tab + 2 * j));  // (cos, sin) of i = 2j, 2j + 1
      const float s0 = cs.y * sign, s1 = cs.w * sign;
      aw[j] = pack_bf16x2(x1.x * cs.x - x2.x * s0, x1.y * cs.z - x2.y * s1);
      bw[j] = pack_bf16x2(x2.x * cs.x + x1.x * s0, x2.y * cs.z + x1.y * s1);
    }
    *reinterpret_cast<uint4*>(lo) = a;
    *reinterpret_cast<uint4*>(hi) = b;
  }
}

// ------------------------------------------------------------------------------------------------ softmax
// One warp per row.  s fp32 (rows, ld) -> p bf16 (rows, ld); fp32 softmax as in x_transformers (dtype=float32).
template <int ITER>
__global__ void __launch_bounds__(256) softmax_fwd_kernel(const float* __restrict__ s, __nv_bfloat16* __restrict__ p, int64_t rows,
                                                          int n_valid, int ld) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* sr = s + row * ld;
  float v[ITER];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < ITER; ++i) {
    const int c = lane + i * 32;
    v[i] = c < n_valid ? __ldg(sr + c) : -INFINITY;
    m = fmaxf(m, v[i]);
  }
  m = warp_max(m);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < ITER; ++i) {
    const int c = lane + i * 32;
    v[i] = c < n_valid ? __expf(v[i] - m) : 0.f;
    sum += v[i];
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  __nv_bfloat16* pr = p + row * ld;
#pragma unroll
  for (int i = 0; i < ITER; ++i) {
    const int c = lane + i * 32;
    if (c < ld) pr[c] = __float2bfloat16(v[i] * inv);
  }
}

// ds = p * (dp - sum_j p_j dp_j) * scale
template <int ITER>
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const __nv_bfloat16* __restrict__ p, const float* __restrict__ dp,
                                                          __nv_bfloat16* __restrict__ ds, float scale, int64_t rows, int n_valid,
                                                          int ld) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const __nv_bfloat16* pr = p + row * ld;
  const float* dpr = dp + row * ld;
  float pv[ITER], dv[ITER];
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < ITER; ++i) {
    const int c = lane + i * 32;
    pv[i] = c < n_valid ? __bfloat162float(pr[c]) : 0.f;
    dv[i] = c < n_valid ? __ldg(dpr + c) : 0.f;
    dot += pv[i] * dv[i];
  }
  dot = warp_sum(dot);
  __nv_bfloat16* dr = ds + row * ld;
#pragma unroll
  for (int i = 0; i < ITER; ++i) {
    const int c = lane + i * 32;
    if (c < ld) dr[c] = __float2bfloat16(pv[i] * (dv[i] - dot) * scale);
  }
}

// ------------------------------------------------------------------------------------------------ column sums
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
  const float4 f = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = f.x, v[1] = f.y, v[2] = f.z, v[3] = f.w;
}
__device__ __forceinline__ void load4(const __nv_bfloat16* p, float (&v)[4]) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
  v[0] = a.x, v[1] = a.y, v[2] = b.x, v[3] = b.y;
}

template <typename TX, typename TY, bool HAS_Y>
__global__ void __launch_bounds__(256) colsum_kernel(const TX* __restrict__ x, const TY* __restrict__ y, float* __restrict__ out,
                                                     int64_t rows, int64_t cols, int64_t ld, int64_t rows_per_block, int vec_ok) {
  // block = 32 (columns, x4 each) x 8 (row lanes); grid.x = column strips of 128, grid.y = row chunks
  __shared__ float red[8][128];
  TRIBE_PDL_ENTRY();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c0 = static_cast<int64_t>(blockIdx.x) * 128 + tx * 4;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  if (c0 < cols) {
    if (vec_ok && c0 + 4 <= cols) {  // one 8/16-byte load per row and operand, 4 rows in flight
      int64_t r = r0 + ty;
      for (; r + 24 < r1; r += 32) {
        float a[4][4], b[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          load4(x + (r + 8 * u) * ld + c0, a[u]);
          if (HAS_Y) load4(y + (r + 8 * u) * ld + c0, b[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[j] += HAS_Y ? a[u][j] * b[u][j] : a[u][j];
      }
      for (; r < r1; r += 8) {
        float a[4], b[4];
        load4(x + r * ld + c0, a);
        if (HAS_Y) load4(y + r * ld + c0, b);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] += HAS_Y ? a[j] * b[j] : a[j];
      }
    } else {
      for (int64_t r = r0 + ty; r < r1; r += 8) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (c0 + j < cols) {
            float v = to_f32<TX>(x[r * ld + c0 + j]);
            if (HAS_Y) v *= to_f32<TY>(y[r * ld + c0 + j]);
            acc[j] += v;
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) red[ty][tx * 4 + j] = acc[j];
  __syncthreads();
  if (threadIdx.x < 128) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
    const int64_t c = static_cast<int64_t>(blockIdx.x) * 128 + threadIdx.x;
    if (c < cols) atomicAdd(out + c, s);
  }
}

// bf16 column sums with 16-byte loads: a thread owns 8 adjacent columns (strip = 256 columns per block), 8 row lanes per
// block, 4 rows in flight per thread (64 B outstanding per thread instead of 32 B with the generic 4-column kernel).
__global__ void __launch_bounds__(256) colsum_bf16x8_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ out, int64_t rows,
                                                            int64_t cols, int64_t ld, int64_t rows_per_block) {
  __shared__ float red[8][256];
  TRIBE_PDL_ENTRY();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t c0 = static_cast<int64_t>(blockIdx.x) * 256 + tx * 8;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (c0 < cols) {  // host guarantees cols % 8 == 0
    auto add = [&](const uint4& q) {
      const float2 a = unpack_bf16x2(q.x), b = unpack_bf16x2(q.y), c = unpack_bf16x2(q.z), d = unpack_bf16x2(q.w);
      acc[0] += a.x, acc[1] += a.y, acc[2] += b.x, acc[3] += b.y, acc[4] += c.x, acc[5] += c.y, acc[6] += d.x, acc[7] += d.y;
    };
    int64_t r = r0 + ty;
    for (; r + 24 < r1; r += 32) {
      uint4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) q[u] = __ldcs(reinterpret_cast<const uint4*>(x + (r + 8 * u) * ld + c0));
#pragma unroll
      for (int u = 0; u < 4; ++u) add(q[u]);
    }
    for (; r < r1; r += 8) add(__ldcs(reinterpret_cast<const uint4*>(x + r * ld + c0)));
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[ty][tx * 8 + j] = acc[j];
  __syncthreads();
  float s_ = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s_ += red[k][threadIdx.x];
  const int64_t c = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (c < cols) atomicAdd(out + c, s_);
}

// ------------------------------------------------------------------------------------------------ casts / axpby
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  const int64_t nvec = n >> 3;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    reinterpret_cast<uint4*>(dst)[i] = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
  }
  for (int64_t i = (nvec << 3) + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = __float2bfloat16(src[i]);
}

__global__ void __launch_bounds__(256) axpby_kernel(const float* __restrict__ src, float* __restrict__ dst, float a, int accumulate,
                                                    int64_t n) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = a * src[i] + (accumulate ? dst[i] : 0.f);
}

// ------------------------------------------------------------------------------------------------ adaptive avg pool
__device__ __forceinline__ int win_start(int i, int t_in, int t_out) { return static_cast<int>((static_cast<int64_t>(i) * t_in) / t_out); }
__device__ __forceinline__ int win_end(int i, int t_in, int t_out) {
  return static_cast<int>((static_cast<int64_t>(i + 1) * t_in + t_out - 1) / t_out);
}

// token-major pooling: x bf16 (B, t_in, C) -> y bf16 (B, t_out, C); 8 channels (16 B) per thread.
__global__ void __launch_bounds__(256) token_pool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int t_in,
                                                             int t_out, int64_t C) {
  const int b = blockIdx.y, i = blockIdx.x;
  const int s = win_start(i, t_in, t_out), e = win_end(i, t_in, t_out);
  const float inv = 1.0f / static_cast<float>(e - s);
  const int64_t nvec = C >> 3;
  for (int64_t v = threadIdx.x; v < nvec; v += blockDim.x) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int t = s; t < e; ++t) {
      const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + (static_cast<int64_t>(b) * t_in + t) * C) + v);
      const float2 a = unpack_bf16x2(u.x), bb = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
      acc[0] += a.x, acc[1] += a.y, acc[2] += bb.x, acc[3] += bb.y, acc[4] += c.x, acc[5] += c.y, acc[6] += d.x, acc[7] += d.y;
    }
    reinterpret_cast<uint4*>(y + (static_cast<int64_t>(b) * t_out + i) * C)[v] =
        make_uint4(pack_bf16x2(acc[0] * inv, acc[1] * inv), pack_bf16x2(acc[2] * inv, acc[3] * inv), pack_bf16x2(acc[4] * inv, acc[5] * inv),
                   pack_bf16x2(acc[6] * inv, acc[7] * inv));
  }
}

template <typename TO>
__device__ __forceinline__ TO from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16(v); }

// bf16 -> bf16 fast path (the training path): 8 channels (16 B) per thread, both source windows of a token loaded before use
__global__ void __launch_bounds__(256) token_pool_bwd_bf16_kernel(const __nv_bfloat16* __restrict__ dy, __nv_bfloat16* __restrict__ dx, int t_in,
                                                                  int t_out, int64_t C) {
  const int b = blockIdx.y, t = blockIdx.x;
  int i_lo = static_cast<int>((static_cast<int64_t>(t) * t_out) / t_in);
  int i_hi = static_cast<int>((static_cast<int64_t>(t + 1) * t_out + t_in - 1) / t_in);
  if (i_hi > t_out) i_hi = t_out;
  const int64_t nvec = C >> 3;
  for (int64_t v = threadIdx.x; v < nvec; v += blockDim.x) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = i_lo; i < i_hi; ++i) {
      const int s = win_start(i, t_in, t_out), e = win_end(i, t_in, t_out);
      if (t >= s && t < e) {
        const float w = 1.0f / static_cast<float>(e - s);
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(dy + (static_cast<int64_t>(b) * t_out + i) * C) + v);
        const float2 a = unpack_bf16x2(u.x), bb = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
        acc[0] += a.x * w, acc[1] += a.y * w, acc[2] += bb.x * w, acc[3] += bb.y * w;
        acc[4] += c.x * w, acc[5] += c.y * w, acc[6] += d.x * w, acc[7] += d.y * w;
      }
    }
    reinterpret_cast<uint4*>(dx + (static_cast<int64_t>(b) * t_in + t) * C)[v] =
        make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]), pack_bf16x2(acc[6], acc[7]));
  }
}

template <typename T, typename TO>
__global__ void __launch_bounds__(256) token_pool_bwd_kernel(const T* __restrict__ dy, TO* __restrict__ dx, int t_in, int t_out, int64_t C) {
  const int b = blockIdx.y, t = blockIdx.x;
  int i_lo = static_cast<int>((static_cast<int64_t>(t) * t_out) / t_in);
  int i_hi = static_cast<int>((static_cast<int64_t>(t + 1) * t_out + t_in - 1) / t_in);
  if (i_hi > t_out) i_hi = t_out;
  for (int64_t c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int i = i_lo; i < i_hi; ++i) {
      const int s = win_start(i, t_in, t_out), e = win_end(i, t_in, t_out);
      if (t >= s && t < e) acc += to_f32<T>(dy[(static_cast<int64_t>(b) * t_out + i) * C + c]) / static_cast<float>(e - s);
    }
    dx[(static_cast<int64_t>(b) * t_in + t) * C + c] = from_f32<TO>(acc);
  }
}

// out[r, c] = (x ? x[r, c] : 0) + pos[r % row_mod, c]   for c in [0, cols)   (positional embedding add, model.py:169-170)
__global__ void __launch_bounds__(256) add_rows_periodic_kernel(const float* __restrict__ x, int64_t ld_x, const float* __restrict__ pos,
                                                                int64_t ld_pos, float* __restrict__ out, int64_t ld_out, int64_t rows,
                                                                int64_t cols, int64_t row_mod) {
  const int64_t total = rows * cols;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    const float base = x ? x[r * ld_x + c] : 0.f;
    out[r * ld_out + c] = base + __ldg(pos + (r % row_mod) * ld_pos + c);
  }
}

// (B, O, T) fp32 -> (B, T, O) bf16
__global__ void __launch_bounds__(256) transpose_cast_bot_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t O,
                                                                 int64_t Tn) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int64_t o0 = static_cast<int64_t>(blockIdx.y) * 32, t0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < 32; r += 8) {
    const int64_t o = o0 + r, t = t0 + lane;
    tile[r][lane] = (o < O && t < Tn) ? __ldg(x + (static_cast<int64_t>(b) * O + o) * Tn + t) : 0.f;
  }
  __syncthreads();
  for (int r = warp; r < 32; r += 8) {
    const int64_t t = t0 + r, o = o0 + lane;
    if (t < Tn && o < O) y[(static_cast<int64_t>(b) * Tn + t) * O + o] = __float2bfloat16(tile[lane][r]);
  }
}

__global__ void __launch_bounds__(256) subject_bias_grad_kernel(const __nv_bfloat16* __restrict__ dy, const long long* __restrict__ subjects,
                                                                float* __restrict__ d_bias, int64_t Tn, int64_t O, int64_t n_subjects) {
  const int b = blockIdx.y;
  const int64_t o = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (o >= O) return;
  const long long s = subjects[b];
  if (s < 0 || s >= n_subjects) return;
  float acc = 0.f;
  for (int64_t t = 0; t < Tn; ++t) acc += __bfloat162float(dy[(static_cast<int64_t>(b) * Tn + t) * O + o]);
  atomicAdd(d_bias + s * O + o, acc);
}

__global__ void check_subjects_kernel(const long long* __restrict__ subjects, int64_t n, int64_t n_subjects, int* __restrict__ flag,
                                      long long* __restrict__ clamped) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const long long s = subjects[i];
    if (s >= n_subjects || s < 0) atomicExch(flag, 1);
    if (clamped) clamped[i] = s < 0 ? 0 : (s >= n_subjects ? n_subjects - 1 : s);
  }
}

}  // namespace tribe

using namespace tribe;

extern "C" int tribe_ingest_features(const void* x, int32_t src_dtype, int64_t B, int64_t L, int64_t D, int64_t T, int32_t layer_mean,
                                     void* out_bf16, int64_t ld_out, int64_t col_off, void* stream) {
  if (!x || !out_bf16 || B <= 0 || L <= 0 || D <= 0 || T <= 0) return set_error(TRIBE_EINVAL, "ingest: bad arguments");
  if (B > 65535) return set_error(TRIBE_EINVAL, "ingest: B > 65535");
  const int64_t feats = layer_mean ? D : L * D;
  dim3 grid(static_cast<unsigned>((T + 63) / 64), static_cast<unsigned>((feats + 63) / 64), static_cast<unsigned>(B));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  switch (src_dtype) {
    case 0: ingest_kernel<float><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(x), L, D, T, layer_mean, o, ld_out, col_off); break;
    case 1: ingest_kernel<double><<<grid, 256, 0, s>>>(reinterpret_cast<const double*>(x), L, D, T, layer_mean, o, ld_out, col_off); break;
    case 2: ingest_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(x), L, D, T, layer_mean, o, ld_out, col_off); break;
    case 3: ingest_kernel<__half><<<grid, 256, 0, s>>>(reinterpret_cast<const __half*>(x), L, D, T, layer_mean, o, ld_out, col_off); break;
    default: return set_error(TRIBE_EINVAL, "ingest: unsupported source dtype");
  }
  TRIBE_CHECK_LAUNCH("ingest");
  return TRIBE_OK;
}

extern "C" int tribe_scalenorm_fwd(const float* x, const float* g, void* y_bf16, float* rnorm, int64_t rows, int64_t dim, float gain_mult, float eps,
                                   void* stream) {
  if (!x || !g || !y_bf16 || rows <= 0 || dim <= 0 || dim % 4) return set_error(TRIBE_EINVAL, "scalenorm_fwd: bad arguments (dim % 4)");
  if (gain_mult < 0.f || eps < 0.f) return set_error(TRIBE_EINVAL, "scalenorm_fwd: gain_mult and eps must be >= 0 (0 = default)");
  const float mult = gain_mult > 0.f ? gain_mult : sqrtf(static_cast<float>(dim));
  const float eps_ = eps > 0.f ? eps : 1e-12f;
  cudaStream_t st_ = reinterpret_cast<cudaStream_t>(stream);
  __nv_bfloat16* yb = reinterpret_cast<__nv_bfloat16*>(y_bf16);
  const bool al = ((reinterpret_cast<uintptr_t>(x) & 15) | (reinterpret_cast<uintptr_t>(y_bf16) & 7)) == 0;
  const int wgrid = grid_for(rows, 8, 148 * 2);
  if (al && dim == 3072) {
    launch_k(scalenorm_fwd_warp_kernel<24>, dim3(wgrid), dim3(256), 0, st_, x, g, yb, rnorm, rows, mult, eps_);
  } else if (al && dim == 384) {
    launch_k(scalenorm_fwd_warp_kernel<3>, dim3(wgrid), dim3(256), 0, st_, x, g, yb, rnorm, rows, mult, eps_);
  } else if (al && dim == 1024) {
    launch_k(scalenorm_fwd_warp_kernel<8>, dim3(wgrid), dim3(256), 0, st_, x, g, yb, rnorm, rows, mult, eps_);
  } else {
    const int grid = grid_for(rows, 1, kMaxBlocks * 4);
    launch_k(scalenorm_fwd_kernel, dim3(grid), dim3(256), 0, st_, x, g, yb, rnorm, rows, static_cast<int>(dim), mult, eps_);
  }
  TRIBE_CHECK_LAUNCH("scalenorm_fwd");
  return TRIBE_OK;
}

extern "C" int tribe_sublayer_bwd(const float* dy_out, const void* d_xn_bf16, const float* x_in, const float* rnorm, const float* g,
                                  const float* rs, float* dx_in, void* dx_in_bf16, float* d_rs, float* d_g, int64_t rows, int64_t dim,
                                  float gain_mult, void* stream) {
  if (gain_mult < 0.f) return set_error(TRIBE_EINVAL, "sublayer_bwd: gain_mult must be >= 0 (0 = sqrt(dim))");
  const float mult = gain_mult > 0.f ? gain_mult : sqrtf(static_cast<float>(dim));
  if (!x_in || rows <= 0 || dim <= 0 || dim % 4 || dim > 4096) return set_error(TRIBE_EINVAL, "sublayer_bwd: bad arguments (dim % 4, dim <= 4096)");
  if (d_xn_bf16 && (!rnorm || !g)) return set_error(TRIBE_EINVAL, "sublayer_bwd: d_xn needs rnorm and g");
  const int nv = static_cast<int>((dim / 4 + 255) / 256);
  static const int sub_bpsm = [] {
    const char* e = getenv("TRIBE_SUBLAYER_BPSM");
    const int v = e ? atoi(e) : 0;
    return v > 0 ? v : 2;
  }();
  const int grid = grid_for(rows, 2, 148 * sub_bpsm);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const __nv_bfloat16* dxn = reinterpret_cast<const __nv_bfloat16*>(d_xn_bf16);
  __nv_bfloat16* dxb = reinterpret_cast<__nv_bfloat16*>(dx_in_bf16);
  const int d = static_cast<int>(dim);
  switch (nv) {
    case 1: launch_k(sublayer_bwd_kernel<1>, dim3(grid), dim3(256), 0, st, dy_out, dxn, x_in, rnorm, g, rs, dx_in, dxb, d_rs, d_g, rows, d, mult); break;
    case 2: launch_k(sublayer_bwd_kernel<2>, dim3(grid), dim3(256), 0, st, dy_out, dxn, x_in, rnorm, g, rs, dx_in, dxb, d_rs, d_g, rows, d, mult); break;
    case 3: launch_k(sublayer_bwd_kernel<3>, dim3(grid), dim3(256), 0, st, dy_out, dxn, x_in, rnorm, g, rs, dx_in, dxb, d_rs, d_g, rows, d, mult); break;
    default: launch_k(sublayer_bwd_kernel<4>, dim3(grid), dim3(256), 0, st, dy_out, dxn, x_in, rnorm, g, rs, dx_in, dxb, d_rs, d_g, rows, d, mult); break;
  }
  TRIBE_CHECK_LAUNCH("sublayer_bwd");
  return TRIBE_OK;
}

extern "C" int tribe_rope_half(void* x_bf16, int64_t rows, int64_t ld, int64_t col_off, int64_t n_heads, int64_t head_dim, int64_t rot_dim,
                               const float* table, int64_t T, float sign, void* stream) {
  if (!x_bf16 || !table || rows <= 0 || n_heads <= 0 || T <= 0 || rot_dim <= 0 || rot_dim > head_dim || rot_dim % 16 || head_dim % 8 || ld % 8 ||
      col_off % 8 || col_off + n_heads * head_dim > ld)
    return set_error(TRIBE_EINVAL, "rope_half: bad arguments (rot_dim % 16, head_dim / ld / col_off % 8, heads inside the row)");
  if ((reinterpret_cast<uintptr_t>(x_bf16) & 15) || (reinterpret_cast<uintptr_t>(table) & 15)) return set_error(TRIBE_EINVAL, "rope_half: 16-byte alignment required");
  const int64_t total = rows * n_heads * (rot_dim / 16);
  rope_half_kernel<<<grid_for(total, 256, kMaxBlocks * 4), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<__nv_bfloat16*>(x_bf16), rows, ld, col_off, static_cast<int>(n_heads), static_cast<int>(head_dim), static_cast<int>(rot_dim),
      reinterpret_cast<const float2*>(table), static_cast<int>(T), sign);
  TRIBE_CHECK_LAUNCH("rope_half");
  return TRIBE_OK;
}

extern "C" int tribe_softmax_fwd(const float* s, void* p_bf16, int64_t rows, int64_t n_valid, int64_t ld, void* stream) {
  if (!s || !p_bf16 || rows <= 0 || n_valid <= 0 || n_valid > ld || ld > 1024) return set_error(TRIBE_EINVAL, "softmax_fwd: bad arguments (ld <= 1024)");
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(p_bf16);
  const int nv = static_cast<int>(n_valid), l = static_cast<int>(ld);
  if (ld <= 128) softmax_fwd_kernel<4><<<grid, 256, 0, st>>>(s, p, rows, nv, l);
  else if (ld <= 320) softmax_fwd_kernel<10><<<grid, 256, 0, st>>>(s, p, rows, nv, l);
  else if (ld <= 512) softmax_fwd_kernel<16><<<grid, 256, 0, st>>>(s, p, rows, nv, l);
  else softmax_fwd_kernel<32><<<grid, 256, 0, st>>>(s, p, rows, nv, l);
  TRIBE_CHECK_LAUNCH("softmax_fwd");
  return TRIBE_OK;
}

extern "C" int tribe_softmax_bwd(const void* p_bf16, const float* dp, void* ds_bf16, float scale, int64_t rows, int64_t n_valid, int64_t ld,
                                 void* stream) {
  if (!p_bf16 || !dp || !ds_bf16 || rows <= 0 || n_valid <= 0 || n_valid > ld || ld > 1024) return set_error(TRIBE_EINVAL, "softmax_bwd: bad arguments");
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(p_bf16);
  __nv_bfloat16* ds = reinterpret_cast<__nv_bfloat16*>(ds_bf16);
  const int nv = static_cast<int>(n_valid), l = static_cast<int>(ld);
  if (ld <= 128) softmax_bwd_kernel<4><<<grid, 256, 0, st>>>(p, dp, ds, scale, rows, nv, l);
  else if (ld <= 320) softmax_bwd_kernel<10><<<grid, 256, 0, st>>>(p, dp, ds, scale, rows, nv, l);
  else if (ld <= 512) softmax_bwd_kernel<16><<<grid, 256, 0, st>>>(p, dp, ds, scale, rows, nv, l);
  else softmax_bwd_kernel<32><<<grid, 256, 0, st>>>(p, dp, ds, scale, rows, nv, l);
  TRIBE_CHECK_LAUNCH("softmax_bwd");
  return TRIBE_OK;
}

extern "C" int tribe_colsum(const void* x, int32_t x_dtype, const void* y, int32_t y_dtype, float* out, int64_t rows, int64_t cols, int64_t ld,
                            int32_t accumulate, void* stream) {
  if (!x || !out || rows <= 0 || cols <= 0 || ld < cols) return set_error(TRIBE_EINVAL, "colsum: bad arguments");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(float) * cols, s);
    if (e != cudaSuccess) return set_cuda_error(e, "colsum memset");
  }
  if (x_dtype == 2 && !y && cols % 8 == 0 && ld % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && cols >= 2048) {
    const int64_t strips8 = (cols + 255) / 256;
    int64_t chunks8 = (148 * 4 + strips8 - 1) / strips8;
    if (chunks8 > (rows + 31) / 32) chunks8 = (rows + 31) / 32;
    if (chunks8 < 1) chunks8 = 1;
    const int64_t rpb8 = (rows + chunks8 - 1) / chunks8;
    dim3 grid8(static_cast<unsigned>(strips8), static_cast<unsigned>((rows + rpb8 - 1) / rpb8));
    launch_k(colsum_bf16x8_kernel, grid8, dim3(256), 0, s, reinterpret_cast<const __nv_bfloat16*>(x), out, rows, cols, ld, rpb8);
    TRIBE_CHECK_LAUNCH("colsum");
    return TRIBE_OK;
  }
  const int64_t strips = (cols + 127) / 128;
  int64_t chunks = (148 * 4 + strips - 1) / strips;
  if (chunks > (rows + 31) / 32) chunks = (rows + 31) / 32;
  if (chunks < 1) chunks = 1;
  const int64_t rpb = (rows + chunks - 1) / chunks;
  dim3 grid(static_cast<unsigned>(strips), static_cast<unsigned>((rows + rpb - 1) / rpb));
  using bf = __nv_bfloat16;
  const int vec_ok = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && (!y || (reinterpret_cast<uintptr_t>(y) & 15) == 0);
  if (x_dtype == 0 && !y) launch_k(colsum_kernel<float, float, false>, grid, dim3(256), 0, s, reinterpret_cast<const float*>(x), nullptr, out, rows, cols, ld, rpb, vec_ok);
  else if (x_dtype == 2 && !y) launch_k(colsum_kernel<bf, bf, false>, grid, dim3(256), 0, s, reinterpret_cast<const bf*>(x), nullptr, out, rows, cols, ld, rpb, vec_ok);
  else if (x_dtype == 0 && y_dtype == 0) launch_k(colsum_kernel<float, float, true>, grid, dim3(256), 0, s, reinterpret_cast<const float*>(x), reinterpret_cast<const float*>(y), out, rows, cols, ld, rpb, vec_ok);
  else if (x_dtype == 2 && y_dtype == 2) launch_k(colsum_kernel<bf, bf, true>, grid, dim3(256), 0, s, reinterpret_cast<const bf*>(x), reinterpret_cast<const bf*>(y), out, rows, cols, ld, rpb, vec_ok);
  else if (x_dtype == 0 && y_dtype == 2) launch_k(colsum_kernel<float, bf, true>, grid, dim3(256), 0, s, reinterpret_cast<const float*>(x), reinterpret_cast<const bf*>(y), out, rows, cols, ld, rpb, vec_ok);
  else if (x_dtype == 2 && y_dtype == 0) launch_k(colsum_kernel<bf, float, true>, grid, dim3(256), 0, s, reinterpret_cast<const bf*>(x), reinterpret_cast<const float*>(y), out, rows, cols, ld, rpb, vec_ok);
  else return set_error(TRIBE_EINVAL, "colsum: unsupported dtype combination");
  TRIBE_CHECK_LAUNCH("colsum");
  return TRIBE_OK;
}

extern "C" int tribe_cast_f32_bf16(const float* src, void* dst_bf16, int64_t n, void* stream) {
  if (!src || !dst_bf16 || n <= 0) return set_error(TRIBE_EINVAL, "cast: bad arguments");
  if ((reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst_bf16) & 15)) return set_error(TRIBE_EINVAL, "cast: 16-byte alignment required");
  cast_f32_bf16_kernel<<<grid_for(n / 8 + 1, 256, kMaxBlocks), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, reinterpret_cast<__nv_bfloat16*>(dst_bf16), n);
  TRIBE_CHECK_LAUNCH("cast_f32_bf16");
  return TRIBE_OK;
}

extern "C" int tribe_axpby_f32(const float* src, float* dst, float a, int32_t accumulate, int64_t n, void* stream) {
  if (!src || !dst || n <= 0) return set_error(TRIBE_EINVAL, "axpby: bad arguments");
  axpby_kernel<<<grid_for(n, 256, kMaxBlocks), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(src, dst, a, accumulate, n);
  TRIBE_CHECK_LAUNCH("axpby");
  return TRIBE_OK;
}

extern "C" int tribe_token_pool_fwd(const void* x_bf16, void* y_bf16, int64_t B, int64_t t_in, int64_t t_out, int64_t C, void* stream) {
  if (!x_bf16 || !y_bf16 || B <= 0 || t_in <= 0 || t_out <= 0 || C <= 0 || C % 8 || B > 65535) return set_error(TRIBE_EINVAL, "token_pool_fwd: bad arguments (C % 8)");
  dim3 grid(static_cast<unsigned>(t_out), static_cast<unsigned>(B));
  token_pool_fwd_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x_bf16), reinterpret_cast<__nv_bfloat16*>(y_bf16), static_cast<int>(t_in), static_cast<int>(t_out), C);
  TRIBE_CHECK_LAUNCH("token_pool_fwd");
  return TRIBE_OK;
}

extern "C" int tribe_token_pool_bwd(const void* dy, int32_t dy_dtype, void* dx, int32_t dx_dtype, int64_t B, int64_t t_in, int64_t t_out,
                                    int64_t C, void* stream) {
  if (!dy || !dx || B <= 0 || t_in <= 0 || t_out <= 0 || C <= 0 || B > 65535) return set_error(TRIBE_EINVAL, "token_pool_bwd: bad arguments");
  dim3 grid(static_cast<unsigned>(t_in), static_cast<unsigned>(B));
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  using bf = __nv_bfloat16;
  const int ti = static_cast<int>(t_in), to = static_cast<int>(t_out);
  if (dy_dtype == 0 && dx_dtype == 0) token_pool_bwd_kernel<float, float><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(dy), reinterpret_cast<float*>(dx), ti, to, C);
  else if (dy_dtype == 2 && dx_dtype == 0) token_pool_bwd_kernel<bf, float><<<grid, 256, 0, s>>>(reinterpret_cast<const bf*>(dy), reinterpret_cast<float*>(dx), ti, to, C);
  else if (dy_dtype == 0 && dx_dtype == 2) token_pool_bwd_kernel<float, bf><<<grid, 256, 0, s>>>(reinterpret_cast<const float*>(dy), reinterpret_cast<bf*>(dx), ti, to, C);
  else if (dy_dtype == 2 && dx_dtype == 2 && C % 8 == 0 && ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0)
    token_pool_bwd_bf16_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const bf*>(dy), reinterpret_cast<bf*>(dx), ti, to, C);
  else if (dy_dtype == 2 && dx_dtype == 2) token_pool_bwd_kernel<bf, bf><<<grid, 256, 0, s>>>(reinterpret_cast<const bf*>(dy), reinterpret_cast<bf*>(dx), ti, to, C);
  else return set_error(TRIBE_EINVAL, "token_pool_bwd: unsupported dtype");
  TRIBE_CHECK_LAUNCH("token_pool_bwd");
  return TRIBE_OK;
}

extern "C" int tribe_add_rows_periodic(const float* x, int64_t ld_x, const float* pos, int64_t ld_pos, float* out, int64_t ld_out, int64_t rows,
                                       int64_t cols, int64_t row_mod, void* stream) {
  if (!pos || !out || rows <= 0 || cols <= 0 || row_mod <= 0) return set_error(TRIBE_EINVAL, "add_rows_periodic: bad arguments");
  add_rows_periodic_kernel<<<grid_for(rows * cols, 256, kMaxBlocks), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, ld_x, pos, ld_pos, out, ld_out,
                                                                                                                 rows, cols, row_mod);
  TRIBE_CHECK_LAUNCH("add_rows_periodic");
  return TRIBE_OK;
}

extern "C" int tribe_transpose_cast_bot(const float* x, void* y_bf16, int64_t B, int64_t O, int64_t T, void* stream) {
  if (!x || !y_bf16 || B <= 0 || O <= 0 || T <= 0 || B > 65535) return set_error(TRIBE_EINVAL, "transpose_cast_bot: bad arguments");
  dim3 grid(static_cast<unsigned>((T + 31) / 32), static_cast<unsigned>((O + 31) / 32), static_cast<unsigned>(B));
  transpose_cast_bot_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, reinterpret_cast<__nv_bfloat16*>(y_bf16), O, T);
  TRIBE_CHECK_LAUNCH("transpose_cast_bot");
  return TRIBE_OK;
}

extern "C" int tribe_subject_bias_grad(const void* dy_bf16, const int64_t* subjects, float* d_bias, int64_t B, int64_t T, int64_t O,
                                       int64_t n_subjects, void* stream) {
  if (!dy_bf16 || !subjects || !d_bias || B <= 0 || T <= 0 || O <= 0 || B > 65535) return set_error(TRIBE_EINVAL, "subject_bias_grad: bad arguments");
  dim3 grid(static_cast<unsigned>((O + 255) / 256), static_cast<unsigned>(B));
  subject_bias_grad_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(dy_bf16), reinterpret_cast<const long long*>(subjects), d_bias, T, O, n_subjects);
  TRIBE_CHECK_LAUNCH("subject_bias_grad");
  return TRIBE_OK;
}

extern "C" int tribe_check_subjects(const int64_t* subjects, int64_t n, int64_t n_subjects, int32_t* flag_out, int64_t* clamped_out, void* stream) {
  if (!subjects || !flag_out || n <= 0 || n_subjects <= 0) return set_error(TRIBE_EINVAL, "check_subjects: bad arguments");
  check_subjects_kernel<<<grid_for(n, 256, 64), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const long long*>(subjects), n, n_subjects, flag_out, reinterpret_cast<long long*>(clamped_out));
  TRIBE_CHECK_LAUNCH("check_subjects");
  return TRIBE_OK;
}
