// B-stationary batched GEMM for the short-K attention contractions (included by gemm_sm100.cu):
//   O = P V,  dV = P^T dO,  dQ = dS K,  dK = dS^T Q      per (window, head): M = 298 rows, N = 384 head dims, K = 298..304.
//
// Why a second kernel.  The generic 1-CTA kernel cuts such a problem into 3 row tiles x 2 column tiles and every tile
// loads its own copy of both operands: 6 x 200 KB = 1.2 MB per (window, head), 154 MB per launch through L2 -> SM for
// 80 MB of operands, and the launch runs at the L2 -> SM rate (profiles/r02_attention_scaling_probe.txt: 3.7 us per tile
// wave for 1 us of tensor work; two CTAs per SM with 128-wide tiles = more copies = slower).  Here one work unit is ALL
// row tiles (MT <= 3) of one 128-column block: per k-block the B sub-tile is loaded once and multiplied with MT A tiles
// into MT accumulators (MT x 128 TMEM columns), so a (window, head) costs 3 x (190 + 80) KB = 0.8 MB (-33 %) and the
// number of work units per launch halves (384 instead of 768: fewer pipeline fills and drains).
//
// Layout per stage: MT A tiles (128 x 64 bf16, 16 KB each) + one B tile (128 x 64, 16 KB); 3 stages.  MT = 3 fills 384
// of the 512 TMEM columns, so the accumulators are single-buffered: the next unit's MMAs wait for the drain while its
// operands already stream into the ring.  Eight epilogue warps (two per TMEM lane quarter, alternate 32-column chunks)
// keep that drain short; bf16 row-major outputs are staged in shared memory and leave through TMA stores.  Operands may
// be K-major or MN-major exactly as in the generic kernel; the fused epilogue is the shared one (bias / rotary /
// residual / transposed stores).  No split-K tail, no K-groups (the host never asks).
#pragma once
#include "gemm_common.cuh"

namespace tribe {

constexpr int kMtBN = 128;
constexpr int kMtThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue
constexpr int kMtStages = 3;

template <int MT>
struct GemmMtCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = kMtBN * BK * 2;
  static constexpr int STAGE_BYTES = MT * A_BYTES + B_BYTES;
  static constexpr int NACC = (2 * MT * kMtBN <= kTmemCols) ? 2 : 1;
  static constexpr int EPI_STAGE_BYTES = 8 * 2 * 2048;  // per epilogue warp: two 32 x 32 bf16 TMA-store tiles (SWIZZLE_64B)
  static constexpr int SMEM_BYTES = kMtStages * STAGE_BYTES + EPI_STAGE_BYTES + 256 + 1024;
};

template <int MT, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kMtThreads, 1) gemm_mt_bf16_kernel(const __grid_constant__ GemmKParams p) {
  using Cfg = GemmMtCfg<MT>;
  constexpr int NACC = Cfg::NACC;
  static_assert(MT >= 2 && MT <= 3, "MT");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* epi_stage = smem + kMtStages * Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_stage + Cfg::EPI_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + kMtStages;
  uint64_t* tfull_bar = empty_bar + kMtStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  asm volatile("griddepcontrol.launch_dependents;");  // PDL: see tribe_internal.h
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tma);
    prefetch_tmap(&p.tmb);
    if (p.tma_store) prefetch_tmap(&p.tmd);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kMtStages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int s = 0; s < 2; ++s) {
        mbar_init(&tfull_bar[s], 1);
        mbar_init(&tempty_bar[s], 256);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_holder, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = blockIdx.x; unit < p.num_tiles; unit += gridDim.x) {
        const TileCoord t = decode_tile(p, unit, kMtBN, MT * BM);
        const int mt_act = min(MT, (p.m - t.m0 + BM - 1) / BM);
        const int a_in = p.a_inner_off + t.zi * p.a_zin_stride;
        const int b_in = p.b_inner_off + t.zi * p.b_zin_stride;
        const int za = batch_coord(p.a_gather, t.z, p.a_zdiv);
        const int zb = batch_coord(p.b_gather, t.z, p.b_zdiv);
        const uint32_t tx = mt_act * Cfg::A_BYTES + Cfg::B_BYTES;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + MT * Cfg::A_BYTES;
          mbar_expect_tx(&full_bar[stage], tx);
          // B first: every MMA of the k-block needs it
          if (!B_MN) {
            tma_load_3d(sb, &p.tmb, &full_bar[stage], b_in + kb * BK, t.n0, zb);
          } else {
#pragma unroll
            for (int j = 0; j < kMtBN / 64; ++j)
              tma_load_3d(sb + j * (BK * 128), &p.tmb, &full_bar[stage], b_in + t.n0 + j * 64, kb * BK, zb);
          }
          for (int i = 0; i < mt_act; ++i) {
            if (!A_MN) {
              tma_load_3d(sa + i * Cfg::A_BYTES, &p.tma, &full_bar[stage], a_in + kb * BK, t.m0 + i * BM, za);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j)
                tma_load_3d(sa + i * Cfg::A_BYTES + j * (BK * 128), &p.tma, &full_bar[stage], a_in + t.m0 + i * BM + j * 64, kb * BK, za);
            }
          }
          if (++stage == kMtStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, kMtBN, A_MN, B_MN);
      const uint32_t a_lbo = A_MN ? p.mn_lbo : p.k_lbo, a_sbo = A_MN ? p.mn_sbo : p.k_sbo;
      const uint32_t b_lbo = B_MN ? p.mn_lbo : p.k_lbo, b_sbo = B_MN ? p.mn_sbo : p.k_sbo;
      constexpr uint32_t a_kstep = A_MN ? 16 * 128 : 32;  // bytes per UMMA_K = 16 along K
      constexpr uint32_t b_kstep = B_MN ? 16 * 128 : 32;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int unit = blockIdx.x; unit < p.num_tiles; unit += gridDim.x) {
        const TileCoord t = decode_tile(p, unit, kMtBN, MT * BM);
        const int mt_act = min(MT, (p.m - t.m0 + BM - 1) / BM);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * (MT * kMtBN);
        uint32_t accumulate = 0;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          // one descriptor per operand and stage; tiles and k-steps are constant increments of its 16-byte address field
          // (60 MMAs of 64 tensor-core cycles per unit: the issuing thread must not spend longer than that on each)
          const uint64_t da0 = make_smem_desc(sa, a_lbo, a_sbo);
          const uint64_t db0 = make_smem_desc(sa + MT * Cfg::A_BYTES, b_lbo, b_sbo);
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t db = db0 + ((kk * b_kstep) >> 4);
#pragma unroll
            for (int i = 0; i < MT; ++i) {
              if (i < mt_act) umma_bf16(d_tmem + i * kMtBN, da0 + ((i * Cfg::A_BYTES + kk * a_kstep) >> 4), db, idesc, accumulate);
            }
            accumulate = 1;
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == kMtStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[acc]);
        if (NACC == 2) {
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1;
        } else {
          acc_phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue: 8 warps, two per TMEM lane quarter
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row_in_tile = q * 32 + lane;
    // bf16 row-major outputs leave through TMA stores (3-D map: columns, rows of one outer batch, outer batch — rows past m
    // are clipped by the hardware): a row-per-thread 16-byte store touches 32 different lines per warp instruction, and
    // that store stream — not the MMAs, not the operand loads — was what a tile of these contractions took its time for.
    const bool scaled = p.alpha != 1.0f;
    uint8_t* my_stage = epi_stage + (warp - 2) * 4096;
    int stage_buf = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int unit = blockIdx.x; unit < p.num_tiles; unit += gridDim.x) {
      const TileCoord t = decode_tile(p, unit, kMtBN, MT * BM);
      const int mt_act = min(MT, (p.m - t.m0 + BM - 1) / BM);
      const long long zoff = static_cast<long long>(t.zo) * p.d_zo + static_cast<long long>(t.zi) * p.d_zi;
      const float* bias = p.bias;
      if (bias && p.bias_gathered) bias += static_cast<long long>(batch_coord(p.b_gather, t.z, p.b_zdiv)) * p.bias_z_stride;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * (MT * kMtBN) + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll 1
      for (int i = 0; i < mt_act; ++i) {
        const int row = t.m0 + i * BM + row_in_tile;
        const bool row_ok = row < p.m;
        const int res_row = p.res_row_mod ? row % p.res_row_mod : row;
        const int pos = p.rope ? row % p.rope_t : 0;
#pragma unroll 1
        for (int c = half; c < kMtBN / 32; c += 2) {
          const int col0 = t.n0 + c * 32;
          if (col0 >= p.n) break;  // warp-uniform
          uint32_t raw[32];
          tmem_ld_32x32(t_addr + i * kMtBN + c * 32, raw);
          tmem_ld_wait();
          float v[32];
          if (scaled) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]) * p.alpha;
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
          }
          if (p.tma_store && col0 + 32 <= p.n) {
            const int row0 = t.m0 + i * BM + q * 32;  // first row of the warp's 32 (warp-uniform)
            if (row0 < p.m) {                         // (a warp whose rows all lie past m stores nothing and consumes no buffer)
              epilogue_math<false>(p, v, row, row_ok, col0, zoff, bias, res_row, pos);
              uint4 pk[4];
#pragma unroll
              for (int j = 0; j < 32; j += 8)
                pk[j >> 3] = make_uint4(pack2(v[j], v[j + 1]), pack2(v[j + 2], v[j + 3]), pack2(v[j + 4], v[j + 5]), pack2(v[j + 6], v[j + 7]));
              uint8_t* sbuf = my_stage + stage_buf * 2048;
              if (lane == 0) tma_store_wait_read1();  // the store issued from this buffer two chunks ago has read it
              __syncwarp();
              stage_write_bf16_sw64(sbuf, lane, pk);
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                tma_store_3d(&p.tmd, sbuf, t.zi * static_cast<int>(p.d_zi) + col0, row0, t.zo);
                tma_store_commit();
              }
              stage_buf ^= 1;
            }
          } else {
            epilogue_chunk<false>(p, v, row, row_ok, col0, zoff, bias, res_row, pos);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      if (NACC == 2) {
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      } else {
        acc_phase ^= 1;
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // bulk stores complete before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace tribe
