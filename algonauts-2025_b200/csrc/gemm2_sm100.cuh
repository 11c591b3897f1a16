// 2-CTA (cta_group::2) variant of the tcgen05 GEMM: a CTA PAIR (cluster of 2 on one TPC) owns a 256 x BN tile.
// Each CTA stages its own 128 rows of A and HALF of the B tile (BN/2 columns); the leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256), the hardware feeds both SMs' tensor cores from both shared memories, and each
// CTA's TMEM holds the accumulator rows of its own half.  Per-SM shared-memory traffic per FLOP drops by a third
// (B is fetched once per pair), which is what limits the 1-CTA 128 x 256 kernel.
//
// Synchronisation (per smem stage / accumulator buffer):
//   full[s]   lives in the LEADER; both CTAs' TMA loads complete_tx on it, the leader arms expect_tx for 2 stages' bytes
//   empty[s]  one per CTA; the leader's tcgen05.commit multicasts the arrive to both
//   tfull[a]  one per CTA (multicast commit) -> each CTA's epilogue drains its own TMEM half
//   tempty[a] lives in the leader, count 512: the epilogue threads of BOTH CTAs arrive on it (remote arrive)
#pragma once
#include "gemm_common.cuh"

namespace tribe {

constexpr int BM2 = 256;  // rows per CTA pair
constexpr int kGemm2Threads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two warps per TMEM lane quarter)
constexpr int kEpi2Threads = 256;
__device__ __forceinline__ void epi2_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all prior MMAs of this thread completed) on the barrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const void* tmap, uint32_t leader_bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(tmap)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

template <int BN, bool ADAM = false>
struct Gemm2Cfg {
  static constexpr int A_BYTES = 128 * BK * 2;        // this CTA's 128 rows of A
  static constexpr int B_BYTES = (BN / 2) * BK * 2;   // this CTA's half of the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // eight per-warp epilogue staging tiles (4 KB each: transpose buffer of the fp32 / optimizer epilogues, TMA-store source
  // of the bf16 ones) sit between the operand ring and the barriers; the ring keeps 6 stages at BN = 256
  static constexpr int EPI_STAGE_BYTES = 8 * kEpiStageFloats * 4;
  static constexpr int BAR_BYTES = 256;
  static constexpr int STAGES_RAW = (227 * 1024 - 1024 - BAR_BYTES - EPI_STAGE_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGE_BYTES + BAR_BYTES + 1024;
};

// EPI: which epilogue family this instance carries — separate instantiations so that each keeps only its own live state
// inside the 168-register budget (with everything in one body the fp32 path spilled its eight in-flight residual
// loads right behind each LDG, serialising them):
//   kEpiGeneric  bf16 outputs through TMA stores (+ every fused epilogue: bias, GELU, GELU', rotary) and the generic fallback
//   kEpiF32      fp32 row-major outputs, STORE / RESIDUAL: transposed, coalesced epilogue
//   kEpiAdam     weight gradients whose epilogue carries the optimizer step (see adam_tile_coalesced)
constexpr int kEpiGeneric = 0, kEpiF32 = 1, kEpiAdam = 2;
template <int BN, bool A_MN, bool B_MN, int EPI = kEpiGeneric>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemm2Threads, 1) gemm2_bf16_kernel(const __grid_constant__ GemmKParams p) {
  constexpr bool ADAM = EPI == kEpiAdam;
  constexpr bool LEAN = EPI != kEpiGeneric;  // only bias / RESIDUAL math can occur
  using Cfg = Gemm2Cfg<BN, ADAM>;
  static_assert(BN == 128 || BN == 256, "2-CTA tiles: BN/2 must be a multiple of 64");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + Cfg::EPI_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tfull_bar = empty_bar + Cfg::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  asm volatile("griddepcontrol.launch_dependents;");  // PDL: see tribe_internal.h

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tma);
    prefetch_tmap(&p.tmb);
    if (p.tma_store) {
      prefetch_tmap(&p.tmd);
      if (p.epilogue == TRIBE_EPI_GELU && p.aux_out) prefetch_tmap(&p.tmaux);
    }
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < Cfg::STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tfull_bar[s], 1);
        mbar_init(&tempty_bar[s], 2 * kEpi2Threads);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc2(tmem_holder, kTmemCols);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // peer barriers are initialised and both halves of the TMEM allocation exist
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  asm volatile("griddepcontrol.wait;" ::: "memory");  // everything above overlapped the previous kernel's tail

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      Work w;
      for (int it = 0; next_work(p, it, w, pair, npairs); ++it) {
        const TileCoord t = decode_tile(p, w.tile, BN, BM2);
        const int m0 = t.m0 + static_cast<int>(rank) * 128;
        const int n0 = t.n0 + static_cast<int>(rank) * (BN / 2);
        const int a_in = p.a_inner_off + t.zi * p.a_zin_stride;
        const int b_in = p.b_inner_off + t.zi * p.b_zin_stride;
        const int za = batch_coord(p.a_gather, t.z, p.a_zdiv);
        const int zb = batch_coord(p.b_gather, t.z, p.b_zdiv);
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          const uint32_t lbar = mapa_shared(smem_u32(&full_bar[stage]), 0);
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
          if (!A_MN) {
            tma_load_3d_2sm(sa, &p.tma, lbar, a_in + kb * BK, m0, za);
          } else {
#pragma unroll
            for (int j = 0; j < 2; ++j) tma_load_3d_2sm(sa + j * (BK * 128), &p.tma, lbar, a_in + m0 + j * 64, kb * BK, za);
          }
          if (!B_MN) {
            tma_load_3d_2sm(sb, &p.tmb, lbar, b_in + kb * BK, n0, zb);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 128; ++j) tma_load_3d_2sm(sb + j * (BK * 128), &p.tmb, lbar, b_in + n0 + j * 64, kb * BK, zb);
          }
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA, single thread)
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM2, BN, A_MN, B_MN);
      const uint32_t a_lbo = A_MN ? p.mn_lbo : p.k_lbo, a_sbo = A_MN ? p.mn_sbo : p.k_sbo;
      const uint32_t b_lbo = B_MN ? p.mn_lbo : p.k_lbo, b_sbo = B_MN ? p.mn_sbo : p.k_sbo;
      constexpr uint32_t a_kstep = A_MN ? 16 * 128 : 32;
      constexpr uint32_t b_kstep = B_MN ? 16 * 128 : 32;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      Work w;
      for (int it = 0; next_work(p, it, w, pair, npairs); ++it) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        uint32_t accumulate = 0;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          // one descriptor per operand and stage; the four k-steps are constant increments of its 16-byte address field, so
          // the issuing thread spends a handful of instructions per MMA (it shares its scheduler with two epilogue warps)
          const uint64_t da0 = make_smem_desc(sa, a_lbo, a_sbo);
          const uint64_t db0 = make_smem_desc(sa + Cfg::A_BYTES, b_lbo, b_sbo);
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            umma2_bf16(d_tmem, da0 + ((kk * a_kstep) >> 4), db0 + ((kk * b_kstep) >> 4), idesc, accumulate);
            accumulate = 1;
          }
          umma2_commit_mc(&empty_bar[stage], 0x3);  // both CTAs' smem slots are free once these MMAs have read them
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma2_commit_mc(&tfull_bar[acc], 0x3);  // accumulator halves complete -> both epilogues
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue (both CTAs, own 128 TMEM lanes)
    const int q = warp & 3;               // TMEM lane quarter
    const int chalf = (warp - 2) >> 2;    // which half of the 32-column chunks this warp takes
    const int row_in_half = q * 32 + lane;
    const uint32_t leader_tempty0 = mapa_shared(smem_u32(&tempty_bar[0]), 0);
    const uint32_t leader_tempty1 = mapa_shared(smem_u32(&tempty_bar[1]), 0);
    float* stage_f = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES) + (warp - 2) * kEpiStageFloats;
    uint8_t* stage_b = reinterpret_cast<uint8_t*>(stage_f);  // bf16 TMA-store tiles: D at +0, aux_out at +2048 (32 rows x 64 B each)
    // where a finished chunk goes (all warp-uniform):
    //   fp32 row-major D, STORE / RESIDUAL   -> transposed through the staging tile, coalesced bias / residual / store
    //   bf16 row-major D with tensor maps     -> fused math in registers, tile staged in SWIZZLE_64B, one TMA store per chunk
    //   weight gradient with an armed optimizer (ADAM instance) -> Adam step through the staging tile
    //   anything else (ragged last chunk, transposed / batched outputs) -> generic row-per-thread epilogue
    const bool path_f32 = EPI == kEpiF32 && p.vec_ok && !p.bias_gathered;  // (host: fp32 row-major D, STORE / RESIDUAL)
    const bool path_tma = EPI == kEpiGeneric && p.tma_store != 0;
    const bool gelu = p.epilogue == TRIBE_EPI_GELU && p.aux_out != nullptr;  // second TMA-store tile only when the pre-activation is wanted
    const bool scaled = p.alpha != 1.0f;
    bool store_pending = false;  // lane 0: a bulk store may still be reading this warp's staging tile
    auto finish_chunk = [&](float (&v)[32], int row, bool row_ok, int col0, long long zoff, const float* bias, int res_row, int pos,
                            const uint4* pre_aux) {
      const bool full_chunk = col0 + 32 <= p.n;
      if constexpr (ADAM) {
        if (p.adam_p && p.vec_ok && full_chunk) {
          epilogue_math<true>(p, v, row, row_ok, col0, zoff, bias, res_row, pos, pre_aux);
          adam_tile_coalesced(p, v, stage_f, lane, row - lane, col0, zoff);
          return;
        }
      }
      if constexpr (EPI == kEpiF32) {
        if (path_f32 && full_chunk) {
          store_tile_f32_coalesced(p, v, stage_f, lane, row - lane, col0, zoff, bias);
          return;
        }
      }
      if constexpr (EPI == kEpiGeneric) {
        if (path_tma && full_chunk) {
          uint4 auxp[4], pk[4];
          epilogue_math<false>(p, v, row, row_ok, col0, zoff, bias, res_row, pos, pre_aux, gelu ? auxp : nullptr);
#pragma unroll
          for (int j = 0; j < 32; j += 8)
            pk[j >> 3] = make_uint4(pack2(v[j], v[j + 1]), pack2(v[j + 2], v[j + 3]), pack2(v[j + 4], v[j + 5]), pack2(v[j + 6], v[j + 7]));
          if (lane == 0 && store_pending) tma_store_wait_read();
          __syncwarp();
          stage_write_bf16_sw64(stage_b, lane, pk);
          if (gelu) stage_write_bf16_sw64(stage_b + 2048, lane, auxp);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&p.tmd, stage_b, col0, row);  // lane 0's row is the tile's first row
            if (gelu) tma_store_2d(&p.tmaux, stage_b + 2048, col0, row);
            tma_store_commit();
          }
          store_pending = true;
          return;
        }
      }
      epilogue_chunk<ADAM, false, LEAN>(p, v, row, row_ok, col0, zoff, bias, res_row, pos, pre_aux);
    };
    int acc = 0;
    uint32_t acc_phase = 0;
    Work w;
    for (int it = 0; next_work(p, it, w, pair, npairs); ++it) {
      const TileCoord t = decode_tile(p, w.tile, BN, BM2);
      const long long zoff = static_cast<long long>(t.zo) * p.d_zo + static_cast<long long>(t.zi) * p.d_zi;
      if constexpr (ADAM) {
        // The optimizer state of this tile (128 rows x BN columns of p, m, v = 3 x 128 KB per CTA) is requested into L2
        // NOW, while the tile's MMAs are still running: 384 row segments of 1 KB, one or two bulk prefetches per epilogue
        // thread.  The epilogue's state loads then are L2 hits instead of HBM round trips (it is latency-bound: each
        // warp has 6 KB in flight per round).
        if (p.adam_p && p.vec_ok) {
          const int ncols = min(BN, p.n - t.n0);
          const uint32_t bytes = static_cast<uint32_t>(ncols) * 4u;
          for (int idx = (warp - 2) * 32 + lane; idx < 3 * 128; idx += kEpi2Threads) {
            const int arr = idx >> 7, prow = t.m0 + static_cast<int>(rank) * 128 + (idx & 127);
            if (prow < p.m && (bytes & 15u) == 0) {
              const float* base = arr == 0 ? p.adam_p : (arr == 1 ? p.adam_m : p.adam_v);
              const float* src = base + zoff + static_cast<long long>(prow) * p.ldd + t.n0;
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
            }
          }
        }
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row = t.m0 + static_cast<int>(rank) * 128 + row_in_half;
      const bool row_ok = row < p.m;
      const float* bias = p.bias;
      if (bias && p.bias_gathered) bias += static_cast<long long>(batch_coord(p.b_gather, t.z, p.b_zdiv)) * p.bias_z_stride;
      const int res_row = p.res_row_mod ? row % p.res_row_mod : row;
      const int pos = p.rope ? row % p.rope_t : 0;
      const uint32_t t_addr = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
      const uint32_t leader_tempty = acc ? leader_tempty1 : leader_tempty0;

      if (!w.partial) {
        // GELU' epilogue (dgrad of FF2): the bf16 pre-activations of the NEXT chunk are requested before this chunk's
        // accumulator is read, one chunk ahead of their use
        const bool pre = !LEAN && p.epilogue == TRIBE_EPI_GELU_BWD && p.vec_ok && row_ok;  // (bf16 outputs only)
        uint4 ax_cur[4], ax_nxt[4];
        auto ld_aux = [&](int c, uint4 (&a)[4]) {
          const uint4* ap = reinterpret_cast<const uint4*>(p.aux_in + static_cast<long long>(row) * p.ld_aux + t.n0 + c * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) a[j] = __ldg(ap + j);
        };
        if (pre && t.n0 + chalf * 32 + 32 <= p.n) ld_aux(chalf, ax_cur);
#pragma unroll 1
        for (int c = chalf; c < BN / 32; c += 2) {
          const int col0 = t.n0 + c * 32;
          if (col0 >= p.n) break;
          const bool have = pre && col0 + 32 <= p.n;
          if (pre && c + 2 < BN / 32 && col0 + 64 + 32 <= p.n) ld_aux(c + 2, ax_nxt);
          uint32_t raw[32];
          tmem_ld_32x32(t_addr + c * 32, raw);
          tmem_ld_wait();
          float v[32];
          if (scaled) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]) * p.alpha;
          } else {  // alpha == 1 (every GEMM of the step but the scaled scores): 32 multiplies per chunk and thread saved
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
          }
          finish_chunk(v, row, row_ok, col0, zoff, bias, res_row, pos, have ? ax_cur : nullptr);
#pragma unroll
          for (int j = 0; j < 4; ++j) ax_cur[j] = ax_nxt[j];
        }
        tc_fence_before();
        mbar_arrive_remote(leader_tempty);
      } else {
        // split-K tail (see gemm_sm100.cu): slot per (tile, slice, CTA rank); each rank reduces its own 128 rows
        const int ti = w.tile - p.full_tiles;
        float* tile_ws = p.ws + (static_cast<size_t>(ti) * 2 + rank) * p.split * (128 * BN);
        float* wrow = tile_ws + static_cast<size_t>(w.slice) * (128 * BN) + static_cast<size_t>(row_in_half) * BN;
#pragma unroll 1
        for (int c = chalf; c < BN / 32; c += 2) {
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
        mbar_arrive_remote(leader_tempty);
        __threadfence();
        epi2_bar_sync();
        int* arrive = p.counters + 4 * ti + 2 * static_cast<int>(rank);
        int* depart = arrive + 1;
        if (warp == 2 && lane == 0) {
          atomicAdd(arrive, 1);
          while (atomicAdd(arrive, 0) < p.split) __nanosleep(64);
          __threadfence();
        }
        epi2_bar_sync();
        const float* rrow = tile_ws + static_cast<size_t>(row_in_half) * BN;
#pragma unroll 1
        for (int c = w.slice + chalf * p.split; c < BN / 32; c += 2 * p.split) {
          const int col0 = t.n0 + c * 32;
          if (col0 >= p.n) break;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.f;
          for (int s2 = 0; s2 < p.split; ++s2) {
            const float4* src = reinterpret_cast<const float4*>(rrow + static_cast<size_t>(s2) * (128 * BN) + c * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 s4 = __ldcg(src + j);
              v[4 * j] += s4.x, v[4 * j + 1] += s4.y, v[4 * j + 2] += s4.z, v[4 * j + 3] += s4.w;
            }
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
          finish_chunk(v, row, row_ok, col0, zoff, bias, res_row, pos, nullptr);
        }
        epi2_bar_sync();
        if (warp == 2 && lane == 0) {
          if (atomicAdd(depart, 1) == p.split - 1) {
            *arrive = 0;
            *depart = 0;
            __threadfence();
          }
        }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0 && store_pending) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // bulk stores complete before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody may free TMEM / exit while the peer can still signal or read
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, kTmemCols);
  }
}

}  // namespace tribe
