// tcgen05 / TMEM / TMA batched GEMM for sm_100a:  D[z] = epilogue(alpha * A[z] * B[z]^T), bf16 in, fp32 accumulate.
//
// One persistent CTA per SM, 192 threads, warp-specialised:
//   warp 0      TMA producer   (cp.async.bulk.tensor.3d into a SWIZZLE_128B smem ring, mbarrier complete_tx)
//   warp 1      MMA issuer     (one thread issues tcgen05.mma 128 x BN x 16 into a double-buffered TMEM accumulator)
//   warps 2..5  epilogue       (tcgen05.ld 32x32b -> registers -> fused epilogue -> global)
// Both operands may be K-major or MN-major (canonical SWIZZLE_128B UMMA layouts), so forward, dgrad and wgrad GEMMs and
// the attention / SubjectLayers contractions all run through this one kernel without materialised transposes.
//
// Scheduling: tiles are dealt round-robin to the persistent CTAs.  When the last wave is ragged (e.g. 456 tiles on 148
// SMs = 3.08 waves for every N = 3072 GEMM of the encoder) the tiles of that partial wave are split along K across
// the otherwise idle CTAs: each CTA stores its K-slice partial into its own fp32 workspace slot, and once all slices
// of a tile have arrived (counter) every slice reduces and finishes a share of the tile's 32-column chunks
// ("data-parallel + split-K tail").
//
// Replaces (reference, relative to /root/reference): projector nn.Linear algonauts2025/model.py:157; the encoder's
// linears and attention einsums behind model.py:173 (x_transformers, restated in oracle/xt_encoder.py); SubjectLayers
// index_select + einsum modeling_utils/modeling_utils/models/common.py:61-66; InfoNCE logits model.py:216.
#include <cuda.h>
#include <cuda_bf16.h>
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <type_traits>
#include <unordered_map>

#include "ptx_sm100.cuh"
#include "tribe_b200.h"
#include "tribe_internal.h"

#include "gemm_common.cuh"
#include "gemm2_sm100.cuh"
#include "gemm_mt_sm100.cuh"

namespace tribe {

template <int BN, bool A_MN, bool B_MN, bool SMALL = false>
__global__ void __launch_bounds__(kGemmThreads, SMALL ? 2 : 1) gemm_bf16_kernel(const __grid_constant__ GemmKParams p) {
  using Cfg = typename std::conditional<SMALL, GemmCfgSmall<BN>, GemmCfg<BN>>::type;
  constexpr uint32_t kCols = SMALL ? 256 : kTmemCols;  // TMEM columns this CTA allocates (two CTAs per SM when SMALL)
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN");
  static_assert(!SMALL || 2 * BN <= 256, "two accumulators must fit half of TMEM");
  static_assert(!B_MN || BN % 64 == 0, "MN-major B tiles are loaded in 64-wide chunks");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tfull_bar = empty_bar + Cfg::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  volatile uint32_t* epi_flag = tmem_holder + 1;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  asm volatile("griddepcontrol.launch_dependents;");  // PDL: see tribe_internal.h

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tma);
    prefetch_tmap(&p.tmb);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < Cfg::STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tfull_bar[s], 1);
        mbar_init(&tempty_bar[s], 128);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_holder, kCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  asm volatile("griddepcontrol.wait;" ::: "memory");  // everything above overlapped the previous kernel's tail
  const int n_kouter = p.kgroup ? p.kgroup_len : 1;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      Work w;
      for (int it = 0; next_work(p, it, w, blockIdx.x, gridDim.x); ++it) {
        const TileCoord t = decode_tile(p, w.tile, BN);
        const int a_in = p.a_inner_off + t.zi * p.a_zin_stride;
        const int b_in = p.b_inner_off + t.zi * p.b_zin_stride;
        for (int ko = 0; ko < n_kouter; ++ko) {
          int za, zb;
          if (p.kgroup) {
            if (static_cast<int>(p.kgroup[ko]) != t.z) continue;
            za = zb = ko;
          } else {
            za = batch_coord(p.a_gather, t.z, p.a_zdiv);
            zb = batch_coord(p.b_gather, t.z, p.b_zdiv);
          }
          for (int kb = w.kb0; kb < w.kb1; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
            uint8_t* sb = sa + Cfg::A_BYTES;
            mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            if (!A_MN) {
              tma_load_3d(sa, &p.tma, &full_bar[stage], a_in + kb * BK, t.m0, za);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                tma_load_3d(sa + j * (BK * 128), &p.tma, &full_bar[stage], a_in + t.m0 + j * 64, kb * BK, za);
            }
            if (!B_MN) {
              tma_load_3d(sb, &p.tmb, &full_bar[stage], b_in + kb * BK, t.n0, zb);
            } else {
#pragma unroll
              for (int j = 0; j < BN / 64; ++j)
                tma_load_3d(sb + j * (BK * 128), &p.tmb, &full_bar[stage], b_in + t.n0 + j * 64, kb * BK, zb);
            }
            if (++stage == Cfg::STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
      const uint32_t a_lbo = A_MN ? p.mn_lbo : p.k_lbo, a_sbo = A_MN ? p.mn_sbo : p.k_sbo;
      const uint32_t b_lbo = B_MN ? p.mn_lbo : p.k_lbo, b_sbo = B_MN ? p.mn_sbo : p.k_sbo;
      constexpr uint32_t a_kstep = A_MN ? 16 * 128 : 32;  // bytes per UMMA_K = 16 along K
      constexpr uint32_t b_kstep = B_MN ? 16 * 128 : 32;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      Work w;
      for (int it = 0; next_work(p, it, w, blockIdx.x, gridDim.x); ++it) {
        const TileCoord t = decode_tile(p, w.tile, BN);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        uint32_t accumulate = 0;
        for (int ko = 0; ko < n_kouter; ++ko) {
          if (p.kgroup && static_cast<int>(p.kgroup[ko]) != t.z) continue;
          for (int kb = w.kb0; kb < w.kb1; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
            const uint64_t da0 = make_smem_desc(sa, a_lbo, a_sbo);  // k-steps = constant increments of the address field
            const uint64_t db0 = make_smem_desc(sa + Cfg::A_BYTES, b_lbo, b_sbo);
#pragma unroll
            for (int kk = 0; kk < BK / 16; ++kk) {
              umma_bf16(d_tmem, da0 + ((kk * a_kstep) >> 4), db0 + ((kk * b_kstep) >> 4), idesc, accumulate);
              accumulate = 1;
            }
            umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
            if (++stage == Cfg::STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (4 warps = 128 TMEM lanes)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row_in_tile = q * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    Work w;
    for (int it = 0; next_work(p, it, w, blockIdx.x, gridDim.x); ++it) {
      const TileCoord t = decode_tile(p, w.tile, BN);
      bool has_k = true;
      if (p.kgroup) {
        has_k = false;
        for (int ko = 0; ko < p.kgroup_len; ++ko) has_k |= (static_cast<int>(p.kgroup[ko]) == t.z);
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row = t.m0 + row_in_tile;
      const bool row_ok = row < p.m;
      const long long zoff = static_cast<long long>(t.zo) * p.d_zo + static_cast<long long>(t.zi) * p.d_zi;
      const float* bias = p.bias;
      if (bias && p.bias_gathered) bias += static_cast<long long>(batch_coord(p.b_gather, t.z, p.b_zdiv)) * p.bias_z_stride;
      const int res_row = p.res_row_mod ? row % p.res_row_mod : row;
      const int pos = p.rope ? row % p.rope_t : 0;
      const uint32_t t_addr = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);

      if (!w.partial) {
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          const int col0 = t.n0 + c * 32;
          if (col0 >= p.n) break;  // warp-uniform
          uint32_t raw[32];
          tmem_ld_32x32(t_addr + c * 32, raw);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = has_k ? __uint_as_float(raw[j]) * p.alpha : 0.0f;
          epilogue_chunk(p, v, row, row_ok, col0, zoff, bias, res_row, pos);
        }
        tc_fence_before();
        mbar_arrive(&tempty_bar[acc]);
      } else {
        // K-slice of a tail tile: reduce into the fp32 workspace, last arriver finishes the tile.
        // Phase 1: this CTA's K-slice partial goes to its own fp32 slot (plain 16-byte stores).
        // Phase 2: once all `split` slices of the tile have landed (arrival counter; every tail unit is resident
        // concurrently: one per CTA, grid <= #SMs), slice s reduces and finishes the 32-column chunks c = s mod split.
        const int ti = w.tile - p.full_tiles;
        const int slice = w.slice;
        float* tile_ws = p.ws + static_cast<size_t>(ti) * p.split * (BM * BN);
        float* wrow = tile_ws + static_cast<size_t>(slice) * (BM * BN) + static_cast<size_t>(row_in_tile) * BN;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          if (t.n0 + c * 32 >= p.n) break;
          uint32_t raw[32];
          tmem_ld_32x32(t_addr + c * 32, raw);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            __stcg(reinterpret_cast<float4*>(wrow + c * 32 + j),
                   make_float4(__uint_as_float(raw[j]), __uint_as_float(raw[j + 1]), __uint_as_float(raw[j + 2]), __uint_as_float(raw[j + 3])));
        }
        tc_fence_before();
        mbar_arrive(&tempty_bar[acc]);
        __threadfence();
        epi_bar_sync();
        int* arrive = p.counters + 4 * ti;
        int* depart = arrive + 1;
        if (warp == 2 && lane == 0) {
          atomicAdd(arrive, 1);
          while (atomicAdd(arrive, 0) < p.split) __nanosleep(64);
          __threadfence();
        }
        epi_bar_sync();
        const float* rrow = tile_ws + static_cast<size_t>(row_in_tile) * BN;
#pragma unroll 1
        for (int c = slice; c < BN / 32; c += p.split) {
          const int col0 = t.n0 + c * 32;
          if (col0 >= p.n) break;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
          for (int s2 = 0; s2 < p.split; ++s2) {
            const float4* src = reinterpret_cast<const float4*>(rrow + static_cast<size_t>(s2) * (BM * BN) + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 s4 = __ldcg(src + j);
              v[4 * j] += s4.x, v[4 * j + 1] += s4.y, v[4 * j + 2] += s4.z, v[4 * j + 3] += s4.w;
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
          epilogue_chunk(p, v, row, row_ok, col0, zoff, bias, res_row, pos);
        }
        epi_bar_sync();
        if (warp == 2 && lane == 0) {
          if (atomicAdd(depart, 1) == p.split - 1) {  // last slice out resets the tile's counters for the next launch
            *arrive = 0;
            *depart = 0;
            __threadfence();
          }
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kCols);
  }
}

// ---------------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  return fn;
}

struct TmapKey {
  const void* ptr;
  int64_t inner, rows, batch, row_stride, batch_stride;
  int box_rows;
  bool operator==(const TmapKey& o) const { return memcmp(this, &o, sizeof(TmapKey)) == 0; }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < sizeof(TmapKey) / 8; ++i) h = (h ^ w[i]) * 1099511628211ull;
    return static_cast<size_t>(h);
  }
};

static int encode_operand(const TribeOperand& op, int box_rows, CUtensorMap* out) {
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  static std::mutex mu;
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = op.ptr, key.inner = op.inner, key.rows = op.rows, key.batch = op.batch, key.row_stride = op.row_stride;
  key.batch_stride = op.batch > 1 ? op.batch_stride : 0, key.box_rows = box_rows;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return TRIBE_OK;
    }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(TRIBE_EDRIVER, "cuTensorMapEncodeTiled entry point not available");
  if ((reinterpret_cast<uintptr_t>(op.ptr) & 15) || (op.row_stride * 2) % 16 || (op.batch > 1 && (op.batch_stride * 2) % 16))
    return set_error(TRIBE_EINVAL, "GEMM operand: pointer and strides must be 16-byte aligned");
  if (op.inner <= 0 || op.rows <= 0 || op.batch <= 0) return set_error(TRIBE_EINVAL, "GEMM operand: empty extent");
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(op.inner), static_cast<cuuint64_t>(op.rows), static_cast<cuuint64_t>(op.batch)};
  cuuint64_t bstride = op.batch > 1 ? static_cast<cuuint64_t>(op.batch_stride) * 2 : static_cast<cuuint64_t>(op.row_stride) * 2 * op.rows;
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(op.row_stride) * 2, bstride};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  alignas(64) CUtensorMap tm;
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(op.ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char msg[200];
    snprintf(msg, sizeof(msg), "cuTensorMapEncodeTiled failed (%d): inner=%lld rows=%lld batch=%lld rs=%lld bs=%lld box=%d",
             static_cast<int>(r), (long long)op.inner, (long long)op.rows, (long long)op.batch, (long long)op.row_stride,
             (long long)op.batch_stride, box_rows);
    return set_error(TRIBE_ETMAP, msg);
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 4096) cache.clear();
    cache[key] = tm;
  }
  *out = tm;
  return TRIBE_OK;
}

// 2-D bf16 output map for the TMA-store epilogue of the 2-CTA kernel: box = 32 columns x 32 rows, SWIZZLE_64B.
static int encode_out_bf16(const void* ptr, int64_t cols, int64_t rows, int64_t ld, CUtensorMap* out) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(TRIBE_EDRIVER, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  alignas(64) CUtensorMap tm;
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(TRIBE_ETMAP, "cuTensorMapEncodeTiled failed for the output map");
  *out = tm;
  return TRIBE_OK;
}

// 3-D bf16 output map of the batched multi-row-tile kernel: (columns, rows of one outer batch, outer batch); rows >= m are
// clipped by the store.  Same 32 x 32 SWIZZLE_64B box as the 2-D map above.
static int encode_out_bf16_3d(const void* ptr, int64_t cols, int64_t rows, int64_t zo, int64_t ld, int64_t zo_stride, CUtensorMap* out) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return set_error(TRIBE_EDRIVER, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(zo)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(zo > 1 ? zo_stride : ld * rows) * 2};
  cuuint32_t box[3] = {32, 32, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  alignas(64) CUtensorMap tm;
  CUresult r = fn(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(TRIBE_ETMAP, "cuTensorMapEncodeTiled failed for the batched output map");
  *out = tm;
  return TRIBE_OK;
}

static std::atomic<int> g_sm_limit{0};  // tribe_gemm_set_sm_limit: 0 = all SMs

static int device_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    // TRIBE_GEMM_SMS=<k>: persistent workers use at most k SMs (leave the rest to concurrently running collectives,
    // whose CTAs would otherwise delay the one-CTA-per-SM grid by a whole kernel)
    if (const char* e = getenv("TRIBE_GEMM_SMS")) {
      const int v = atoi(e);
      if (v >= 2 && v < n) n = v & ~1;
    }
  }
  return n;
}

static int num_sms() {
  const int n = device_sms(), lim = g_sm_limit.load(std::memory_order_relaxed);
  return (lim >= 2 && lim < n) ? (lim & ~1) : n;
}

extern "C" int tribe_gemm_set_sm_limit(int32_t n_sms) {
  if (n_sms < 0) return set_error(TRIBE_EINVAL, "gemm_set_sm_limit: negative");
  g_sm_limit.store(n_sms, std::memory_order_relaxed);
  return TRIBE_OK;
}

template <int BN, bool A_MN, bool B_MN, bool SMALL = false>
static int launch_gemm(const GemmKParams& kp, int grid, cudaStream_t stream) {
  using Cfg = typename std::conditional<SMALL, GemmCfgSmall<BN>, GemmCfg<BN>>::type;
  static bool attr_set = false;
  auto kern = gemm_bf16_kernel<BN, A_MN, B_MN, SMALL>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(gemm)");
    attr_set = true;
  }
  cudaError_t e = launch_k(kern, dim3(grid), dim3(kGemmThreads), Cfg::SMEM_BYTES, stream, kp);
  count_launch();
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "gemm launch");
  return TRIBE_OK;
}

template <int BN, bool A_MN, bool B_MN, int EPI = kEpiGeneric>
static int launch_gemm2(const GemmKParams& kp, int grid, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BN, EPI == kEpiAdam>;
  static bool attr_set = false;
  auto kern = gemm2_bf16_kernel<BN, A_MN, B_MN, EPI>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(gemm2)");
    attr_set = true;
  }
  cudaError_t e = launch_k(kern, dim3(grid), dim3(kGemm2Threads), Cfg::SMEM_BYTES, stream, kp);  // __cluster_dims__(2,1,1): grid is a multiple of 2
  count_launch();
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "gemm2 launch");
  return TRIBE_OK;
}

template <int MT, bool A_MN, bool B_MN>
static int launch_gemm_mt(const GemmKParams& kp, int grid, cudaStream_t stream) {
  using Cfg = GemmMtCfg<MT>;
  static bool attr_set = false;
  auto kern = gemm_mt_bf16_kernel<MT, A_MN, B_MN>;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(gemm_mt)");
    attr_set = true;
  }
  cudaError_t e = launch_k(kern, dim3(grid), dim3(kMtThreads), Cfg::SMEM_BYTES, stream, kp);
  count_launch();
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "gemm_mt launch");
  return TRIBE_OK;
}

template <int MT>
static int dispatch_major_mt(const GemmKParams& kp, int grid, bool a_mn, bool b_mn, cudaStream_t s) {
  if (!a_mn && !b_mn) return launch_gemm_mt<MT, false, false>(kp, grid, s);
  if (a_mn && !b_mn) return launch_gemm_mt<MT, true, false>(kp, grid, s);
  if (!a_mn && b_mn) return launch_gemm_mt<MT, false, true>(kp, grid, s);
  return launch_gemm_mt<MT, true, true>(kp, grid, s);
}

template <int BN, int EPI>
static int dispatch_major2e(const GemmKParams& kp, int grid, bool a_mn, bool b_mn, cudaStream_t s) {
  if (!a_mn && !b_mn) return launch_gemm2<BN, false, false, EPI>(kp, grid, s);
  if (a_mn && !b_mn) return launch_gemm2<BN, true, false, EPI>(kp, grid, s);
  if (!a_mn && b_mn) return launch_gemm2<BN, false, true, EPI>(kp, grid, s);
  return launch_gemm2<BN, true, true, EPI>(kp, grid, s);
}

template <int BN>
static int dispatch_major2(const GemmKParams& kp, int grid, bool a_mn, bool b_mn, cudaStream_t s) {
  if (kp.adam_p) return launch_gemm2<BN, true, true, kEpiAdam>(kp, grid, s);  // (host: wgrad form only)
  if (kp.f32_coalesce) return dispatch_major2e<BN, kEpiF32>(kp, grid, a_mn, b_mn, s);
  return dispatch_major2e<BN, kEpiGeneric>(kp, grid, a_mn, b_mn, s);
}

template <int BN>
static int dispatch_major(const GemmKParams& kp, int grid, bool a_mn, bool b_mn, cudaStream_t s) {
  if (!a_mn && !b_mn) return launch_gemm<BN, false, false>(kp, grid, s);
  if (a_mn && !b_mn) return launch_gemm<BN, true, false>(kp, grid, s);
  if constexpr (BN % 64 == 0) {
    if (!a_mn && b_mn) return launch_gemm<BN, false, true>(kp, grid, s);
    return launch_gemm<BN, true, true>(kp, grid, s);
  } else {
    return set_error(TRIBE_EINVAL, "block_n=160 requires a K-major B operand");
  }
}

static int dispatch_major_small(const GemmKParams& kp, int grid, bool a_mn, bool b_mn, cudaStream_t s) {
  if (!a_mn && !b_mn) return launch_gemm<128, false, false, true>(kp, grid, s);
  if (a_mn && !b_mn) return launch_gemm<128, true, false, true>(kp, grid, s);
  if (!a_mn && b_mn) return launch_gemm<128, false, true, true>(kp, grid, s);
  return launch_gemm<128, true, true, true>(kp, grid, s);
}

static int pick_block_n(const TribeGemm& g) {
  if (g.block_n) return g.block_n;
  const bool b_mn = g.b.mn_major != 0;
  if (g.n <= 128) return 128;
  if (g.n <= 160 && !b_mn) return 160;
  if (g.n <= 192) return 192;
  if (g.n <= 256) return 256;
  if (g.n <= 320 && !b_mn) return 160;
  if (g.n <= 384) return 192;
  return 256;
}

static int gemm_impl(const TribeGemm* g, void* stream, uint32_t k_lbo, uint32_t k_sbo, uint32_t mn_lbo, uint32_t mn_sbo) {
  if (!g || !g->a.ptr || !g->b.ptr || !g->d) return set_error(TRIBE_EINVAL, "gemm: null pointer");
  if (g->m <= 0 || g->n <= 0 || g->k <= 0 || g->batch <= 0) return set_error(TRIBE_EINVAL, "gemm: empty problem");
  int bn = pick_block_n(*g);
  // short-K batched problems (attention P.V, dV, dQ, dK: 5-6 k-blocks per tile): 128-wide tiles, two CTAs per SM.
  // Measured on B200 (profiles/r01_attention_microbench.txt): P.V 35.8 us vs 32.8 us with the 192-wide single-CTA tiles and
  // no difference in the train step, so this stays opt-in (TRIBE_GEMM_SMALL=1).
  static const int allow_small = [] {
    const char* e = getenv("TRIBE_GEMM_SMALL");
    return e ? atoi(e) : 0;
  }();
  const bool small = allow_small && g->block_n == 0 && !g->kgroup && (g->k + BK - 1) / BK <= 8 && (g->n % 128 == 0 || g->n <= 128) &&
                     static_cast<int64_t>(g->batch) * ((g->m + BM - 1) / BM) * ((g->n + 127) / 128) >= 2 * num_sms();
  if (small) bn = 128;
  // batched short-K problems with 2-3 row tiles and N a multiple of 128 (attention P.V, dV, dQ, dK): B-stationary units of
  // all row tiles x 128 columns (gemm_mt_sm100.cuh): a third less L2 -> SM traffic, half the work units.  TRIBE_GEMM_MT=0: off.
  static const int allow_mt = [] {
    const char* e = getenv("TRIBE_GEMM_MT");
    return e ? atoi(e) : 1;
  }();
  const int m_tiles = (g->m + BM - 1) / BM;
  const bool use_mt = allow_mt && !small && g->block_n == 0 && !g->kgroup && !g->adam_p && g->batch >= 2 && (m_tiles == 2 || m_tiles == 3) &&
                      (g->k + BK - 1) / BK <= 8 && g->n % 128 == 0;
  if (use_mt) bn = 128;
  if (bn != 128 && bn != 160 && bn != 192 && bn != 256) return set_error(TRIBE_EINVAL, "gemm: block_n must be 128/160/192/256");
  if (g->epilogue == TRIBE_EPI_ROPE) {
    if (!g->rope || g->rope_t <= 0 || g->head_dim % 32 || g->rope_dim % 32 || g->rope_dim > g->head_dim)
      return set_error(TRIBE_EINVAL, "gemm: rope epilogue needs a table and head_dim/rope_dim multiples of 32");
  }
  if (g->epilogue == TRIBE_EPI_GELU_BWD && !g->aux_in) return set_error(TRIBE_EINVAL, "gemm: GELU_BWD epilogue needs aux_in");
  if (g->epilogue == TRIBE_EPI_RESIDUAL && !g->res) return set_error(TRIBE_EINVAL, "gemm: RESIDUAL epilogue needs res");

  GemmKParams kp;
  memset(&kp, 0, sizeof(kp));
  const bool a_mn = g->a.mn_major != 0, b_mn = g->b.mn_major != 0;
  // CTA pairs (cta_group::2, 256 x BN tiles) for the big GEMMs; single CTAs for short-M / grouped problems.
  static const int allow_2cta = [] {
    const char* e = getenv("TRIBE_GEMM_2CTA");
    return e ? atoi(e) : 1;
  }();
  bool use2 = allow_2cta && !use_mt && bn == 256 && !g->kgroup && g->m >= 1024 && (num_sms() % 2 == 0);
  if (use2 && g->adam_p && !(a_mn && b_mn)) use2 = false;  // only the wgrad form (both operands MN-major) has a 2-CTA Adam instance
  int rc = encode_operand(g->a, a_mn ? BK : BM, &kp.tma);
  if (rc) return rc;
  rc = encode_operand(g->b, b_mn ? BK : (use2 ? bn / 2 : bn), &kp.tmb);
  if (rc) return rc;
  kp.m = g->m, kp.n = g->n, kp.k = g->k, kp.batch = g->batch, kp.z_inner = g->z_inner > 0 ? g->z_inner : 1;
  kp.a_inner_off = g->a.inner_off, kp.a_zin_stride = g->a.zin_stride, kp.a_zdiv = g->a.zdiv > 0 ? g->a.zdiv : 1;
  kp.b_inner_off = g->b.inner_off, kp.b_zin_stride = g->b.zin_stride, kp.b_zdiv = g->b.zdiv > 0 ? g->b.zdiv : 1;
  kp.a_gather = reinterpret_cast<const long long*>(g->a.gather);
  kp.b_gather = reinterpret_cast<const long long*>(g->b.gather);
  kp.kgroup = reinterpret_cast<const long long*>(g->kgroup), kp.kgroup_len = g->kgroup_len;
  kp.d = g->d, kp.d_f32 = g->d_f32, kp.d_transposed = g->d_transposed;
  kp.ldd = g->ldd, kp.d_zo = g->d_zo_stride, kp.d_zi = g->d_zi_stride;
  kp.epilogue = g->epilogue, kp.alpha = g->alpha;
  kp.bias = g->bias, kp.bias_gathered = g->bias_gathered, kp.bias_z_stride = g->bias_z_stride;
  kp.res = g->res, kp.ld_res = g->ld_res, kp.res_row_mod = g->res_row_mod, kp.res_batched = g->res_batched, kp.rscale = g->rscale;
  kp.aux_in = reinterpret_cast<const __nv_bfloat16*>(g->aux_in);
  kp.aux_out = reinterpret_cast<__nv_bfloat16*>(g->aux_out), kp.ld_aux = g->ld_aux;
  kp.rope = reinterpret_cast<const float2*>(g->rope);
  kp.rope_t = g->rope_t, kp.rope_dim = g->rope_dim, kp.head_dim = g->head_dim > 0 ? g->head_dim : 32;
  kp.rope_cols = g->rope_cols, kp.rope_sign = g->rope_sign;
  kp.k_lbo = k_lbo ? k_lbo : 16, kp.k_sbo = k_sbo ? k_sbo : 1024;
  kp.mn_lbo = mn_lbo ? mn_lbo : BK * 128, kp.mn_sbo = mn_sbo ? mn_sbo : 1024;
  const int bm = use2 ? BM2 : (use_mt ? m_tiles * BM : BM);
  kp.m_blocks = (g->m + bm - 1) / bm, kp.n_blocks = (g->n + bn - 1) / bn;
  kp.num_tiles = kp.m_blocks * kp.n_blocks * g->batch, kp.num_kb = (g->k + BK - 1) / BK;
  // Tile order: every wave of ~#SM tiles streams one operand completely and re-reads the other; re-read the smaller
  // one so that it stays L2-resident (126 MB) — e.g. FF2 (A = 117 MB activations, B = 75 MB weights) goes N-fastest.
  kp.raster_n_fast = (static_cast<double>(g->m) > static_cast<double>(g->n)) ? 1 : 0;
  // 16-byte vector epilogue accesses need aligned bases / leading dimensions.
  const int esz = g->d_f32 ? 4 : 2;
  auto al16 = [](const void* p_) { return (reinterpret_cast<uintptr_t>(p_) & 15) == 0; };
  bool vec = al16(g->d) && (g->ldd * esz) % 16 == 0 && (g->d_zo_stride * esz) % 16 == 0 && (g->d_zi_stride * esz) % 16 == 0;
  if (g->bias) vec = vec && al16(g->bias) && (g->bias_z_stride * 4) % 16 == 0;
  if (g->res) vec = vec && al16(g->res) && (g->ld_res * 4) % 16 == 0;
  if (g->rscale) vec = vec && al16(g->rscale);
  if (g->aux_in) vec = vec && al16(g->aux_in) && (g->ld_aux * 2) % 16 == 0;
  if (g->aux_out) vec = vec && al16(g->aux_out) && (g->ld_aux * 2) % 16 == 0;
  if (g->rope && !al16(g->rope)) return set_error(TRIBE_EINVAL, "gemm: rope table must be 16-byte aligned");
  if (g->adam_p) {
    if (!g->adam_m || !g->adam_v || !g->adam_hyper) return set_error(TRIBE_EINVAL, "gemm: fused Adam needs adam_m, adam_v and adam_hyper");
    if (!g->d_f32 || g->d_transposed) return set_error(TRIBE_EINVAL, "gemm: fused Adam needs a row-major fp32 output");
    if (g->epilogue != TRIBE_EPI_STORE && g->epilogue != TRIBE_EPI_RESIDUAL)
      return set_error(TRIBE_EINVAL, "gemm: fused Adam combines with the STORE and RESIDUAL (accumulate) epilogues only");
    if (!al16(g->adam_hyper)) return set_error(TRIBE_EINVAL, "gemm: adam_hyper must be 16-byte aligned");
    vec = vec && al16(g->adam_p) && al16(g->adam_m) && al16(g->adam_v) && al16(g->adam_shadow);
    kp.adam_p = g->adam_p, kp.adam_m = g->adam_m, kp.adam_v = g->adam_v;
    kp.adam_shadow = reinterpret_cast<__nv_bfloat16*>(g->adam_shadow), kp.adam_hyper = g->adam_hyper, kp.adam_keep_grad = g->adam_keep_grad;
  }
  kp.vec_ok = vec ? 1 : 0;
  // bf16 row-major outputs of the 2-CTA kernel leave through TMA stores (TRIBE_TMA_STORE=0: per-thread 16-byte stores)
  static const int allow_tma_store = [] {
    const char* e = getenv("TRIBE_TMA_STORE");
    return e ? atoi(e) : 1;
  }();
  static const int allow_f32_coalesce = [] {
    const char* e = getenv("TRIBE_EPI_COALESCE");
    return e ? atoi(e) : 1;
  }();
  kp.f32_coalesce = allow_f32_coalesce && use2 && g->d_f32 && !g->d_transposed && !g->adam_p &&
                    (g->epilogue == TRIBE_EPI_STORE || g->epilogue == TRIBE_EPI_RESIDUAL);
  if (allow_tma_store && use2 && vec && !g->d_f32 && !g->d_transposed && g->batch == 1 && !g->adam_p && g->ldd % 8 == 0) {
    rc = encode_out_bf16(g->d, g->n, g->m, g->ldd, &kp.tmd);
    if (rc) return rc;
    if (g->epilogue == TRIBE_EPI_GELU && g->aux_out) {
      rc = encode_out_bf16(g->aux_out, g->n, g->m, g->ld_aux, &kp.tmaux);
      if (rc) return rc;
    }
    kp.tma_store = 1;
  }
  if (allow_tma_store && use_mt && vec && !g->d_f32 && !g->d_transposed && g->ldd % 8 == 0) {
    // D(zo, zi, row, col) = d + zo * d_zo + zi * d_zi + row * ldd + col: the inner batch index must be a column offset
    const int zin = kp.z_inner, zout = g->batch / zin;
    const int64_t cols = zin > 1 ? static_cast<int64_t>(zin - 1) * g->d_zi_stride + g->n : g->n;
    if (g->batch % zin == 0 && cols <= g->ldd && (zin == 1 || g->d_zi_stride >= g->n) && (zout == 1 || g->d_zo_stride >= static_cast<int64_t>(g->m) * g->ldd)) {
      rc = encode_out_bf16_3d(g->d, cols, g->m, zout, g->ldd, g->d_zo_stride, &kp.tmd);
      if (rc) return rc;
      kp.tma_store = 1;
    }
  }

  // ---- schedule: whole tiles round-robin; the ragged last wave is split along K when a workspace is provided
  const int workers_max = use2 ? num_sms() / 2 : (small ? 2 * num_sms() : num_sms());  // persistent CTAs, or CTA pairs
  const int grid = kp.num_tiles < workers_max ? kp.num_tiles : workers_max;
  kp.full_tiles = kp.num_tiles, kp.tail_units = 0, kp.split = 1, kp.kb_per = kp.num_kb;
  // (measured on B200: the tail split pays off for deep contractions — +15..18 % at K >= 9216 — and is neutral to
  //  slightly negative at K = 3072, where a tile is only ~27 us long; hence the K >= 6144 gate.)
  static const int split_gate_kb = [] {
    const char* e = getenv("TRIBE_SPLITK_GATE_KB");
    const int v = e ? atoi(e) : 0;
    return v > 0 ? v : 96;
  }();
  if (g->splitk_ws && !g->kgroup && kp.num_tiles > grid && kp.num_kb >= (use2 ? split_gate_kb / 2 : split_gate_kb) && al16(g->splitk_ws)) {
    const int full = (kp.num_tiles / grid) * grid;
    const int rem = kp.num_tiles - full;
    if (rem > 0 && rem <= kMaxTailTiles) {
      static const int min_kb = [] {
        const char* e = getenv("TRIBE_SPLITK_MIN_KB");
        const int v = e ? atoi(e) : 0;
        return v > 0 ? v : 8;
      }();
      int split = grid / rem;
      if (split > kp.num_kb / min_kb) split = kp.num_kb / min_kb;  // keep >= min_kb K-blocks per slice
      if (split > bn / 32) split = bn / 32;                        // phase 2 hands out 32-column chunks
      const size_t head = static_cast<size_t>(kMaxTailTiles) * 4 * sizeof(int);
      const size_t per_slot = static_cast<size_t>(bm) * bn * sizeof(float);
      while (split >= 2 && head + static_cast<size_t>(rem) * split * per_slot > static_cast<size_t>(g->splitk_ws_bytes)) --split;
      if (split >= 2) {
        kp.kb_per = (kp.num_kb + split - 1) / split;
        kp.split = (kp.num_kb + kp.kb_per - 1) / kp.kb_per;
        kp.full_tiles = full;
        kp.tail_units = rem * kp.split;
        kp.counters = reinterpret_cast<int*>(g->splitk_ws);
        kp.ws = reinterpret_cast<float*>(reinterpret_cast<char*>(g->splitk_ws) + head);
      }
    }
  }

  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (use_mt) return m_tiles == 3 ? dispatch_major_mt<3>(kp, grid, a_mn, b_mn, s) : dispatch_major_mt<2>(kp, grid, a_mn, b_mn, s);
  if (use2) return dispatch_major2<256>(kp, 2 * grid, a_mn, b_mn, s);
  if (small) return dispatch_major_small(kp, grid, a_mn, b_mn, s);
  switch (bn) {
    case 128: return dispatch_major<128>(kp, grid, a_mn, b_mn, s);
    case 160: return dispatch_major<160>(kp, grid, a_mn, b_mn, s);
    case 192: return dispatch_major<192>(kp, grid, a_mn, b_mn, s);
    default: return dispatch_major<256>(kp, grid, a_mn, b_mn, s);
  }
}

}  // namespace tribe

extern "C" int tribe_gemm_bf16(const TribeGemm* g, void* stream) { return tribe::gemm_impl(g, stream, 0, 0, 0, 0); }

extern "C" int tribe_gemm_bf16_probe(const TribeGemm* g, void* stream, uint32_t k_lbo, uint32_t k_sbo, uint32_t mn_lbo,
                                     uint32_t mn_sbo) {
  return tribe::gemm_impl(g, stream, k_lbo, k_sbo, mn_lbo, mn_sbo);
}

#include "attn_sm100.cuh"
