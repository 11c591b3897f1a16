// Shared device-side pieces of the tcgen05 GEMM kernels (1-CTA: gemm_sm100.cu, 2-CTA pairs: gemm2_sm100.cuh):
// kernel parameter block, tile decoding, persistent work scheduling (whole tiles + split-K tail) and the fused epilogue.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "adam_math.cuh"
#include "ptx_sm100.cuh"
#include "tribe_b200.h"

namespace tribe {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kGemmThreads = 192;
constexpr int kTmemCols = 512;
constexpr int kMaxTailTiles = 160;  // split-K counters (4 ints per tail tile) reserved at the head of the workspace

struct alignas(64) GemmKParams {
  CUtensorMap tma, tmb;
  CUtensorMap tmd, tmaux;  // TMA-store maps of D and aux_out (bf16 outputs of the 2-CTA kernel, box 32 x 32, SWIZZLE_64B)
  int tma_store;           // 1: tmd (and tmaux for the GELU epilogue) are valid
  int f32_coalesce;        // 1: fp32 row-major outputs of the 2-CTA kernel take the transposed (coalesced) epilogue
  int m, n, k, batch, z_inner;
  int a_inner_off, a_zin_stride, a_zdiv;
  int b_inner_off, b_zin_stride, b_zdiv;
  const long long* a_gather;
  const long long* b_gather;
  const long long* kgroup;
  int kgroup_len;
  void* d;
  int d_f32, d_transposed, vec_ok;
  long long ldd, d_zo, d_zi;
  int epilogue;
  float alpha;
  const float* bias;
  int bias_gathered;
  long long bias_z_stride;
  const float* res;
  long long ld_res;
  int res_row_mod, res_batched;
  const float* rscale;
  const __nv_bfloat16* aux_in;
  __nv_bfloat16* aux_out;
  long long ld_aux;
  const float2* rope;
  int rope_t, rope_dim, head_dim, rope_cols;
  float rope_sign;
  uint32_t k_lbo, k_sbo, mn_lbo, mn_sbo;
  int m_blocks, n_blocks, num_tiles, num_kb, raster_n_fast;
  // split-K tail
  int full_tiles, tail_units, split, kb_per;
  float* ws;
  int* counters;
  // optimizer step fused into the epilogue (weight gradients): same element offsets as D
  float* adam_p;
  float* adam_m;
  float* adam_v;
  __nv_bfloat16* adam_shadow;
  const float* adam_hyper;
  int adam_keep_grad;
};

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES_RAW = (220 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024: manual 1 KiB alignment
};

// Short-K problems (the attention contractions: 6 k-blocks per tile) are bound by per-tile pipeline latency, not by
// bandwidth: a 3-stage ring (96 KiB) and half of TMEM (2 x 128 columns) let TWO CTAs share an SM and overlap each other's
// fill / drain phases.
template <int BN>
struct GemmCfgSmall {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = 3;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;
};

// Phi(-|x|) = erfc(|x| / sqrt 2) / 2 = exp(-x^2 / 2) * poly(t), t = 1 / (1 + p |x| / sqrt 2)  (Abramowitz-Stegun 7.1.26 with the
// coefficients halved): 11 instructions instead of erff()'s ~25, |error| <= 1e-7 absolute; also returns e = exp(-x^2 / 2).
__device__ __forceinline__ float normal_tail(float x, float& e) {
  float t;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * -0.72134752044448170f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(fabsf(x), 0.23164189815521240f, 1.0f)));
  float poly = 0.5307027145f;  // a5 / 2 ... a1 / 2
  poly = fmaf(poly, t, -0.7265760135f);
  poly = fmaf(poly, t, 0.7107068705f);
  poly = fmaf(poly, t, -0.142248368f);
  poly = fmaf(poly, t, 0.127414796f);
  return poly * t * e;
}
// GELU(x) = x Phi(x) (exact-erf GELU of the FF block, oracle/xt_encoder.py; |error| <= 5e-7 absolute on [-10, 10] against
// float64, the result is rounded to bf16).  TRIBE_EXACT_GELU (compile-time) restores the erff forms.
__device__ __forceinline__ float gelu_erf(float x) {
#ifdef TRIBE_EXACT_GELU
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
#else
  float e;
  const float tail = normal_tail(x, e);
  return x * (x >= 0.f ? 1.0f - tail : tail);
#endif
}
// GELU'(x) = Phi(x) + x phi(x) for the dgrad-FF2 epilogue (12288 x 4768 evaluations per layer and step).  erff() + __expf() are
// ~36 instructions per element; here ONE exponential e = exp(-x^2/2) serves both terms (phi = e / sqrt(2 pi), Phi from
// normal_tail): 16 instructions, |error| <= 3e-7 absolute (checked against float64 on [-8, 8]; the factor multiplies a
// bf16-rounded gradient).  Measured: dgrad-FF2 294.8 -> 273.2 us (tools/gemm_bench.py).
__device__ __forceinline__ float gelu_erf_grad(float x) {
#ifdef TRIBE_EXACT_GELU
  return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
#else
  float e;
  const float tail = normal_tail(x, e);
  const float Phi = x >= 0.f ? 1.0f - tail : tail;
  return fmaf(x * 0.3989422804014327f, e, Phi);
#endif
}

struct TileCoord {
  int z, zi, zo, m0, n0;
};

__device__ __forceinline__ TileCoord decode_tile(const GemmKParams& p, int tile, int bn, int bm = BM) {
  TileCoord t;
  int mb, nb, rest;
  if (p.raster_n_fast) {  // consecutive tiles (one wave) share an A row-block and sweep B: B is the re-read operand
    nb = tile % p.n_blocks;
    rest = tile / p.n_blocks;
    mb = rest % p.m_blocks;
    t.z = rest / p.m_blocks;
  } else {                // consecutive tiles share a B column-block and sweep A: A is the re-read operand
    mb = tile % p.m_blocks;
    rest = tile / p.m_blocks;
    nb = rest % p.n_blocks;
    t.z = rest / p.n_blocks;
  }
  t.zi = t.z % p.z_inner;
  t.zo = t.z / p.z_inner;
  t.m0 = mb * bm;
  t.n0 = nb * bn;
  return t;
}

__device__ __forceinline__ int batch_coord(const long long* gather, int z, int zdiv) {
  int zz = z / zdiv;
  return gather ? static_cast<int>(gather[zz]) : zz;
}

// One unit of work of a persistent CTA: a whole tile, or one K-slice of a tile of the ragged last wave.
struct Work {
  int tile, kb0, kb1, slice;
  bool partial;
};

// worker = persistent CTA (1-CTA kernel) or CTA pair (2-CTA kernel); nworkers = how many of them the grid holds.
__device__ __forceinline__ bool next_work(const GemmKParams& p, int it, Work& w, int worker, int nworkers) {
  const int idx = worker + it * nworkers;
  if (idx < p.full_tiles) {
    w.tile = idx, w.kb0 = 0, w.kb1 = p.num_kb, w.slice = 0, w.partial = false;
    return true;
  }
  const int u = idx - p.full_tiles;
  if (u >= p.tail_units) return false;
  w.tile = p.full_tiles + u / p.split;
  w.slice = u % p.split;
  w.kb0 = w.slice * p.kb_per;
  w.kb1 = min(p.num_kb, w.kb0 + p.kb_per);
  w.partial = true;
  return true;
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Adam step on 32 consecutive elements of one weight row whose finished gradient sits in registers (wgrad GEMMs):
// 12 B read + 14 B written per parameter, all streaming (nothing is re-read before the next forward pass); the
// gradient itself never has to reach memory.  A thread owns a whole 128-byte line of each state array per chunk.
template <bool VEC = true>
__device__ __forceinline__ void adam_chunk(const GemmKParams& p, const float (&g)[32], long long off, int nvalid, bool full) {
  const float4 h0 = __ldg(reinterpret_cast<const float4*>(p.adam_hyper));
  const float2 h1 = __ldg(reinterpret_cast<const float2*>(p.adam_hyper + 4));
  const float beta1 = h0.x, beta2 = h0.y, step_size = h0.z, inv_bc2_sqrt = h0.w, eps = h1.x, wd = h1.y;
  float* pp = p.adam_p + off;
  float* pm = p.adam_m + off;
  float* pv = p.adam_v + off;
  __nv_bfloat16* ps = p.adam_shadow ? p.adam_shadow + off : nullptr;
  if (VEC && full) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float4 P[4], M[4], V[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        P[j] = __ldcs(reinterpret_cast<const float4*>(pp) + half * 4 + j);
        M[j] = __ldcs(reinterpret_cast<const float4*>(pm) + half * 4 + j);
        V[j] = __ldcs(reinterpret_cast<const float4*>(pv) + half * 4 + j);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = half * 16 + j * 4;
        adam_one(P[j].x, g[c], M[j].x, V[j].x, beta1, beta2, step_size, inv_bc2_sqrt, eps, wd);
        adam_one(P[j].y, g[c + 1], M[j].y, V[j].y, beta1, beta2, step_size, inv_bc2_sqrt, eps, wd);
        adam_one(P[j].z, g[c + 2], M[j].z, V[j].z, beta1, beta2, step_size, inv_bc2_sqrt, eps, wd);
        adam_one(P[j].w, g[c + 3], M[j].w, V[j].w, beta1, beta2, step_size, inv_bc2_sqrt, eps, wd);
        __stcs(reinterpret_cast<float4*>(pp) + half * 4 + j, P[j]);
        __stcs(reinterpret_cast<float4*>(pm) + half * 4 + j, M[j]);
        __stcs(reinterpret_cast<float4*>(pv) + half * 4 + j, V[j]);
      }
      if (ps) {
#pragma unroll
        for (int j = 0; j < 4; j += 2)
          __stcs(reinterpret_cast<uint4*>(ps) + half * 2 + (j >> 1),
                 make_uint4(pack2(P[j].x, P[j].y), pack2(P[j].z, P[j].w), pack2(P[j + 1].x, P[j + 1].y), pack2(P[j + 1].z, P[j + 1].w)));
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {  // fully unrolled: g[] must stay in registers
      if (j < nvalid) {
        float P = pp[j], M = pm[j], V = pv[j];
        adam_one(P, g[j], M, V, beta1, beta2, step_size, inv_bc2_sqrt, eps, wd);
        pp[j] = P, pm[j] = M, pv[j] = V;
        if (ps) ps[j] = __float2bfloat16(P);
      }
    }
  }
}

// Fused epilogue of 32 consecutive columns of one output row (v already scaled by alpha).
// pre_aux (optional): the chunk's 32 bf16 aux_in values (GELU' operand) already loaded by the caller — issued BEFORE the
// TMEM load so that the global-load latency is not exposed in the epilogue (ncu: 15 % of the dgrad-FF2 samples sat on the
// first use of this load).
// Part 1: everything that changes the VALUES of the chunk (bias, activation, residual, rotary); aux_out side store of GELU.
// LEAN: only what a weight-gradient GEMM can ask for (bias, RESIDUAL accumulation) is compiled in.
// aux_pack (optional, GELU): receives the chunk's 32 bf16 pre-activations instead of the direct aux_out store.
template <bool LEAN = false>
__device__ __forceinline__ void epilogue_math(const GemmKParams& p, float (&v)[32], int row, bool row_ok, int col0, long long zoff,
                                              const float* bias, int res_row, int pos, const uint4* pre_aux = nullptr, uint4* aux_pack = nullptr) {
  const int nvalid = min(32, p.n - col0);
  const bool full = (nvalid == 32) && p.vec_ok;

  if (bias) {
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + col0 + j));
        v[j] += b4.x, v[j + 1] += b4.y, v[j + 2] += b4.z, v[j + 3] += b4.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) v[j] += __ldg(bias + col0 + j);
    }
  }

  if (!LEAN && p.epilogue == TRIBE_EPI_GELU) {
    if (aux_pack) {
#pragma unroll
      for (int j = 0; j < 32; j += 8)
        aux_pack[j >> 3] = make_uint4(pack2(v[j], v[j + 1]), pack2(v[j + 2], v[j + 3]), pack2(v[j + 4], v[j + 5]), pack2(v[j + 6], v[j + 7]));
    } else if (row_ok && p.aux_out) {  // (aux_out == NULL: inference, nobody will read the pre-activation)
      __nv_bfloat16* ap = p.aux_out + static_cast<long long>(row) * p.ld_aux + col0;
      if (full) {
#pragma unroll
        for (int j = 0; j < 32; j += 8)
          *reinterpret_cast<uint4*>(ap + j) = make_uint4(pack2(v[j], v[j + 1]), pack2(v[j + 2], v[j + 3]), pack2(v[j + 4], v[j + 5]), pack2(v[j + 6], v[j + 7]));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) ap[j] = __float2bfloat16(v[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
  } else if (!LEAN && p.epilogue == TRIBE_EPI_GELU_BWD) {
    if (row_ok) {
      const __nv_bfloat16* ap = p.aux_in + static_cast<long long>(row) * p.ld_aux + col0;
      if (full) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const uint4 pk = pre_aux ? pre_aux[j >> 3] : __ldg(reinterpret_cast<const uint4*>(ap + j));
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __bfloat1622float2(h[e]);
            v[j + 2 * e] *= gelu_erf_grad(f.x);
            v[j + 2 * e + 1] *= gelu_erf_grad(f.y);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) v[j] *= gelu_erf_grad(__bfloat162float(ap[j]));
      }
    }
  } else if (p.epilogue == TRIBE_EPI_RESIDUAL) {
    if (row_ok) {
      const float* rp = p.res + (p.res_batched ? zoff : 0) + static_cast<long long>(res_row) * p.ld_res + col0;
      if (full) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 r4 = __ldg(reinterpret_cast<const float4*>(rp + j));
          float4 s4 = make_float4(1.f, 1.f, 1.f, 1.f);
          if (p.rscale) s4 = __ldg(reinterpret_cast<const float4*>(p.rscale + col0 + j));
          v[j] += r4.x * s4.x, v[j + 1] += r4.y * s4.y, v[j + 2] += r4.z * s4.z, v[j + 3] += r4.w * s4.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) v[j] += __ldg(rp + j) * (p.rscale ? __ldg(p.rscale + col0 + j) : 1.0f);
      }
    }
  } else if (!LEAN && p.epilogue == TRIBE_EPI_ROPE) {
    const int cih = col0 % p.head_dim;  // chunk-uniform: head_dim, rope_dim are multiples of 32
    if (col0 < p.rope_cols && cih < p.rope_dim) {
      const float2* tab = p.rope + static_cast<long long>(pos) * (p.rope_dim >> 1) + (cih >> 1);
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const float4 cs = __ldg(reinterpret_cast<const float4*>(tab + j));  // (cos, sin) of two pairs
        const float s0 = cs.y * p.rope_sign, s1 = cs.w * p.rope_sign;
        const float x0 = v[2 * j], x1 = v[2 * j + 1], y0 = v[2 * j + 2], y1 = v[2 * j + 3];
        v[2 * j] = x0 * cs.x - x1 * s0;
        v[2 * j + 1] = x1 * cs.x + x0 * s0;
        v[2 * j + 2] = y0 * cs.z - y1 * s1;
        v[2 * j + 3] = y1 * cs.z + y0 * s1;
      }
    }
  }

}

// Part 2: where the chunk goes (optimizer step and / or the D store).
// ADAM_VEC = false: the caller handles full chunks itself (adam_tile_coalesced) and only ragged ones arrive here.
template <bool ADAM = true, bool ADAM_VEC = true>
__device__ __forceinline__ void epilogue_store(const GemmKParams& p, float (&v)[32], int row, bool row_ok, int col0, long long zoff) {
  const int nvalid = min(32, p.n - col0);
  const bool full = (nvalid == 32) && p.vec_ok;
  if (!row_ok) return;
  if constexpr (ADAM) {
    if (p.adam_p) {  // host side guarantees d_f32 && !d_transposed
      adam_chunk<ADAM_VEC>(p, v, zoff + static_cast<long long>(row) * p.ldd + col0, nvalid, full);
      if (!p.adam_keep_grad) return;
    }
  }
  if (p.d_transposed) {
    if (p.d_f32) {
      float* dp = reinterpret_cast<float*>(p.d) + zoff + static_cast<long long>(col0) * p.ldd + row;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) dp[static_cast<long long>(j) * p.ldd] = v[j];
    } else {
      __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(p.d) + zoff + static_cast<long long>(col0) * p.ldd + row;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) dp[static_cast<long long>(j) * p.ldd] = __float2bfloat16(v[j]);
    }
  } else if (p.d_f32) {
    float* dp = reinterpret_cast<float*>(p.d) + zoff + static_cast<long long>(row) * p.ldd + col0;
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) dp[j] = v[j];
    }
  } else {
    __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(p.d) + zoff + static_cast<long long>(row) * p.ldd + col0;
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 8)
        *reinterpret_cast<uint4*>(dp + j) = make_uint4(pack2(v[j], v[j + 1]), pack2(v[j + 2], v[j + 3]), pack2(v[j + 4], v[j + 5]), pack2(v[j + 6], v[j + 7]));
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < nvalid) dp[j] = __float2bfloat16(v[j]);
    }
  }
}

template <bool ADAM = true, bool ADAM_VEC = true, bool LEAN = false>
__device__ __forceinline__ void epilogue_chunk(const GemmKParams& p, float (&v)[32], int row, bool row_ok, int col0, long long zoff,
                                               const float* bias, int res_row, int pos, const uint4* pre_aux = nullptr) {
  epilogue_math<LEAN>(p, v, row, row_ok, col0, zoff, bias, res_row, pos, pre_aux);
  epilogue_store<ADAM, ADAM_VEC>(p, v, row, row_ok, col0, zoff);
}

// Adam step on a warp's 32-row x 32-column gradient tile with COALESCED state traffic: the tile (thread = row after
// tcgen05.ld) goes through a padded shared-memory buffer (36-float rows: conflict-free 16-byte writes by row and reads by
// quarter-row), after which lane l owns columns 4*(l%8)..+3 of rows l/8, l/8 + 4, ...: every warp-wide access covers four
// full 128-byte lines of p / m / v (row-per-thread accesses touch 32 lines for the same bytes and ran the fused step at
// a third of the stand-alone kernel's bandwidth).  Four row groups (12 x 16-byte loads per lane) are in flight at once.
// ---- per-warp epilogue staging tile (2-CTA kernel): 32 rows x 32 fp32 columns = 4 KB, 16-byte chunk c of row r stored at
// chunk (c ^ (r & 7)): conflict-free both for the writers (thread = row, after tcgen05.ld) and for the readers
// (lane = 4 columns of rows l/8, l/8 + 4, ...), who then touch global memory in full 128-byte lines.
constexpr int kEpiStageFloats = 32 * 32;
__device__ __forceinline__ void stage_write_row(float* stage, int lane, const float (&g)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(stage + lane * 32 + ((j ^ (lane & 7)) << 2)) = make_float4(g[4 * j], g[4 * j + 1], g[4 * j + 2], g[4 * j + 3]);
}
__device__ __forceinline__ float4 stage_read(const float* stage, int r, int c4) {
  return *reinterpret_cast<const float4*>(stage + r * 32 + ((c4 ^ (r & 7)) << 2));
}

// Adam step on a warp's 32-row x 32-column gradient tile with COALESCED state traffic (row-per-thread accesses touch 32
// lines for the same bytes and ran the fused step at a third of the stand-alone kernel's bandwidth).  Four row groups
// (12 x 16-byte loads per lane) are in flight at once.
__device__ __forceinline__ void adam_tile_coalesced(const GemmKParams& p, const float (&g)[32], float* stage, int lane, int row0, int col0,
                                                    long long zoff) {
  __syncwarp();  // the previous chunk's reads of `stage` are complete
  stage_write_row(stage, lane, g);
  __syncwarp();
  const float4 h0 = __ldg(reinterpret_cast<const float4*>(p.adam_hyper));
  const float2 h1 = __ldg(reinterpret_cast<const float2*>(p.adam_hyper + 4));
  const float beta1 = h0.x, beta2 = h0.y, step_size = h0.z, inv_bc2_sqrt = h0.w, eps = h1.x, wd = h1.y;
  const int sub = lane >> 3, c4 = lane & 7;
#pragma unroll
  for (int it0 = 0; it0 < 8; it0 += 4) {
    float4 P[4], M[4], V[4];
    long long off[4];
    bool ok[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = (it0 + u) * 4 + sub;
      ok[u] = row0 + r < p.m;
      off[u] = zoff + static_cast<long long>(row0 + r) * p.ldd + col0 + c4 * 4;
      if (ok[u]) {
        P[u] = __ldcs(reinterpret_cast<const float4*>(p.adam_p + off[u]));
        M[u] = __ldcs(reinterpret_cast<const float4*>(p.adam_m + off[u]));
        V[u] = __ldcs(reinterpret_cast<const float4*>(p.adam_v + off[u]));
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (!ok[u]) continue;
      const float4 G = stage_read(stage, (it0 + u) * 4 + sub, c4);
      adam_one(P[u].x, G.x, M[u].x, V[u].x, beta1, beta2, step_size, inv_bc2_sqrt, eps, wd);
      adam_one(P[u].y, G.y, M[u].y, V[u].y, beta1, beta2, step_size, inv_bc2_sqrt, eps, wd);
      adam_one(P[u].z, G.z, M[u].z, V[u].z, beta1, beta2, step_size, inv_bc2_sqrt, eps, wd);
      adam_one(P[u].w, G.w, M[u].w, V[u].w, beta1, beta2, step_size, inv_bc2_sqrt, eps, wd);
      __stcs(reinterpret_cast<float4*>(p.adam_p + off[u]), P[u]);
      __stcs(reinterpret_cast<float4*>(p.adam_m + off[u]), M[u]);
      __stcs(reinterpret_cast<float4*>(p.adam_v + off[u]), V[u]);
      if (p.adam_shadow) __stcs(reinterpret_cast<uint2*>(p.adam_shadow + off[u]), make_uint2(pack2(P[u].x, P[u].y), pack2(P[u].z, P[u].w)));
      if (p.adam_keep_grad) __stcs(reinterpret_cast<float4*>(reinterpret_cast<float*>(p.d) + off[u]), G);
    }
  }
}

// fp32 outputs (projector / out-projection / FF2 forward with their residual, weight gradients): the accumulator chunk
// (thread = row) is transposed through the staging tile; bias, residual * rscale and the store then run with lane = 4
// columns, i.e. every warp-wide access covers four full 128-byte lines of res / D (row-per-thread accesses need 32
// line visits for the same bytes, which kept the LSU busier than the tensor pipe on the K = 3072 GEMMs).  All eight
// residual loads of the chunk are issued before the first use.
__device__ __forceinline__ void store_tile_f32_coalesced(const GemmKParams& p, const float (&acc)[32], float* stage, int lane, int row0, int col0,
                                                         long long zoff, const float* bias) {
  __syncwarp();
  stage_write_row(stage, lane, acc);
  __syncwarp();
  const int sub = lane >> 3, c4 = lane & 7;
  const int col = col0 + c4 * 4;
  float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), s4 = make_float4(1.f, 1.f, 1.f, 1.f);
  if (bias) b4 = __ldg(reinterpret_cast<const float4*>(bias + col));
  const bool has_res = p.epilogue == TRIBE_EPI_RESIDUAL;
  if (has_res && p.rscale) s4 = __ldg(reinterpret_cast<const float4*>(p.rscale + col));
  float4 R[8];
  if (has_res) {
    // residual row of output row r: r % res_row_mod (positional-embedding table) or r; one modulo per chunk, then stepped
    const int mod = p.res_row_mod;
    int rr = row0 + sub;
    if (mod) rr %= mod;
    const float* rbase = p.res + (p.res_batched ? zoff : 0) + col;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      if (row0 + it * 4 + sub < p.m) R[it] = __ldg(reinterpret_cast<const float4*>(rbase + static_cast<long long>(rr) * p.ld_res));
      rr += 4;
      if (mod) {
        while (rr >= mod) rr -= mod;
      }
    }
  }
  float* dbase = reinterpret_cast<float*>(p.d) + zoff + col;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + sub, row = row0 + r;
    if (row >= p.m) continue;
    float4 o = stage_read(stage, r, c4);
    o.x += b4.x, o.y += b4.y, o.z += b4.z, o.w += b4.w;
    if (has_res) o.x += R[it].x * s4.x, o.y += R[it].y * s4.y, o.z += R[it].z * s4.z, o.w += R[it].w * s4.w;
    *reinterpret_cast<float4*>(dbase + static_cast<long long>(row) * p.ldd) = o;
  }
}

// bf16 outputs through a TMA store: the warp's 32 x 32 chunk goes to its staging tile in the SWIZZLE_64B layout the tensor
// map expects (64-byte rows; 16-byte chunk c of row r at chunk c ^ ((r >> 1) & 3): conflict-free for thread = row
// writers), is published to the async proxy and stored by one lane with cp.async.bulk.tensor (rows past m are clipped
// by the hardware).  `pk` = the thread's 32 packed bf16 values.
__device__ __forceinline__ void stage_write_bf16_sw64(uint8_t* stage, int lane, const uint4 (&pk)[4]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(stage + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) = pk[c];
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

}  // namespace tribe
