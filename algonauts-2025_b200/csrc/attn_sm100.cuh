// Attention scores with the row softmax fused into the tcgen05 epilogue (included at the end of gemm_sm100.cu).
//
// x_transformers attention (restated in oracle/xt_encoder.py; reached through algonauts2025/model.py:173) computes
//   sim = q k^T * d^-1/2 ;  attn = softmax(sim, dim=-1, dtype=float32) ;  out = attn v
// and its backward needs  dS = P o (dP - rowsum(dP o P)) * d^-1/2  with  dP = dO v^T.
// TRIBE's sequences are 298 tokens long, so a WHOLE score row (<= 320 keys) fits in one CTA's TMEM accumulator
// (128 lanes x 320 fp32 columns of the 512): the fp32 scores never travel to HBM.  Compared with the generic path
// (GEMM -> fp32 S in HBM -> softmax kernel -> bf16 P) this removes 2 x 46 MB of traffic and one launch per layer and
// direction.
//
//   mode 0 (forward):   out = P  = softmax(scale * A B^T)            A = q rows, B = k rows (both K-major over head dims)
//   mode 1 (backward):  out = dS = P o (A B^T - rowsum(A B^T o P)) * scale      A = dO rows, B = v rows, P read back (bf16)
//
// One persistent CTA per SM, 320 threads: warp 0 = TMA producer (3-stage ring; A box 64 x 128, B as one or two
// 64 x 160 boxes — key rows >= T are zero-filled by TMA), warp 1 = one thread issuing tcgen05.mma 128 x 160 x 16 per B
// part into TMEM columns [0,160) / [160,320), warps 2..9 = epilogue, TWO threads per query row (each owns half of the
// row's columns; maxima / sums / dot products meet in a 2 KB exchange area), three (mode 0) or two (mode 1) sweeps over
// the row's TMEM columns.  (Round 1 ran four epilogue warps: forward 57 -> 46 us with eight.)  The accumulator is single-buffered (2 x 320 columns do not fit), so MMA
// and epilogue of one CTA alternate while the producer already streams the next tile's operands.

namespace tribe {

constexpr int kAttnNP = 160;                          // UMMA N of one B part
constexpr int kAttnAB = BM * BK * 2;                  // 16 KiB
constexpr int kAttnBB = kAttnNP * BK * 2;             // 20 KiB per part
constexpr int kAttnStage = kAttnAB + 2 * kAttnBB;     // 56 KiB
constexpr int kAttnStages = 3;
constexpr int kAttnThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)
constexpr int kAttnEpiStage = 8 * 2 * 2048;  // per epilogue warp: two 32 x 32 bf16 TMA-store tiles (SWIZZLE_64B)
constexpr int kAttnSmem = kAttnStages * kAttnStage + 3072 + kAttnEpiStage + 1024;  // 3072: barriers (256) + exchange area (2048), padded

struct alignas(64) AttnKParams {
  CUtensorMap tma, tmb;
  CUtensorMap tmo;  // 3-D store map of out: (Tp columns, T rows, batch x heads), box 32 x 32, SWIZZLE_64B
  int tma_store;
  int T, Tp, heads, dh, n_parts, num_kb, m_blocks, num_tiles;
  int a_off, b_off;
  float scale;
  int mode;
  const __nv_bfloat16* p_in;
  __nv_bfloat16* out;
  uint32_t k_lbo, k_sbo;
};

__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

// 2^x as ONE MUFU.EX2 (exp2f() wraps it in a range test and two conditional multiplies to produce denormal results: four
// instructions per score, twice per score in the forward epilogue).  Probabilities below 2^-126 become exact zeros.
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint4 attn_pack8(const float* v) {
  return make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
}

__global__ void __launch_bounds__(kAttnThreads, 1) attn_scores_kernel(const __grid_constant__ AttnKParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kAttnStages * kAttnStage);
  uint64_t* empty_bar = full_bar + kAttnStages;
  uint64_t* tfull_bar = empty_bar + kAttnStages;
  uint64_t* tempty_bar = tfull_bar + 1;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty_bar + 1);
  float* xch = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 256);  // 2 x (2 halves x 128 rows) floats
  uint8_t* epi_stage = reinterpret_cast<uint8_t*>(full_bar) + 3072;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  asm volatile("griddepcontrol.launch_dependents;");  // PDL: see tribe_internal.h
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tma);
    prefetch_tmap(&p.tmb);
    if (p.tma_store) prefetch_tmap(&p.tmo);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kAttnStages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(tfull_bar, 1);
      mbar_init(tempty_bar, 256);
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
  const uint32_t stage_tx = kAttnAB + p.n_parts * kAttnBB;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int z = tile / p.m_blocks, mb = tile - z * p.m_blocks;
        const int b = z / p.heads, h = z - b * p.heads;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kAttnStage;
          uint8_t* sb = sa + kAttnAB;
          mbar_expect_tx(&full_bar[stage], stage_tx);
          tma_load_3d(sa, &p.tma, &full_bar[stage], p.a_off + h * p.dh + kb * BK, mb * BM, b);
          for (int j = 0; j < p.n_parts; ++j)
            tma_load_3d(sb + j * kAttnBB, &p.tmb, &full_bar[stage], p.b_off + h * p.dh + kb * BK, j * kAttnNP, b);
          if (++stage == kAttnStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, kAttnNP, false, false);
      const bool two_parts = p.n_parts > 1;
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar, acc_phase ^ 1);
        tc_fence_after();
        uint32_t accumulate = 0;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kAttnStage);
          // one descriptor per operand and stage; k-steps and the second B part are constant increments of its 16-byte
          // address field (the issuing thread must not spend longer on an MMA than the tensor core does: 80 cycles here)
          const uint64_t da0 = make_smem_desc(sa, p.k_lbo, p.k_sbo);
          const uint64_t db0 = make_smem_desc(sa + kAttnAB, p.k_lbo, p.k_sbo);
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            umma_bf16(tmem_base, da0 + ((kk * 32) >> 4), db0 + ((kk * 32) >> 4), idesc, accumulate);
            if (two_parts) umma_bf16(tmem_base + kAttnNP, da0 + ((kk * 32) >> 4), db0 + ((kAttnBB + kk * 32) >> 4), idesc, accumulate);
            accumulate = 1;
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == kAttnStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(tfull_bar);
        acc_phase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // Eight epilogue warps: two per TMEM lane quarter, each thread owns HALF of its query row's score columns, so two
    // warps per scheduler overlap each other's TMEM-load latencies; row maxima / sums / dot products are combined through
    // a 2 KB exchange area (one named barrier per exchange).
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row_in_tile = q * 32 + lane;
    const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const int nchunks = p.n_parts * (kAttnNP / 32);
    const int c_mid = (nchunks + 1) >> 1;
    const int c_lo = half ? c_mid : 0, c_hi = half ? nchunks : c_mid;  // this thread's 32-column chunks
    const float sl2 = p.scale * 1.4426950408889634f;  // exp(scale * x) = exp2(sl2 * x)
    float* xch_a = xch, *xch_b = xch + 256;
    auto epi_sync = [] { asm volatile("bar.sync 2, 256;" ::: "memory"); };
    auto exchange = [&](float* area, float mine) {  // returns the partner thread's value
      area[half * 128 + row_in_tile] = mine;
      epi_sync();
      return area[(half ^ 1) * 128 + row_in_tile];
    };
    // Output chunks (the warp's 32 rows x 32 columns) leave through a TMA store from a per-warp staging tile: the
    // row-per-thread 16-byte stores they replace touched 32 different lines per warp instruction (4864 of them per tile).
    uint8_t* my_stage = epi_stage + (warp - 2) * 4096;
    int stage_buf = 0;
    auto store_chunk = [&](const float* v, int col0, int row0, int z) {  // warp-uniform arguments except v
      if (row0 >= p.T) return;  // the warp's 32 rows all lie past T: nothing to store (and no buffer is consumed)
      uint4 pk[4];
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) pk[g8] = attn_pack8(v + g8 * 8);
      uint8_t* sbuf = my_stage + stage_buf * 2048;
      if (lane == 0) tma_store_wait_read1();  // the store issued from this buffer two chunks ago has read it
      __syncwarp();
      stage_write_bf16_sw64(sbuf, lane, pk);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        tma_store_3d(&p.tmo, sbuf, col0, row0, z);  // rows >= T and columns >= Tp are clipped by the hardware
        tma_store_commit();
      }
      stage_buf ^= 1;
    };
    uint32_t acc_phase = 0;
    int tile_parity = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, tile_parity ^= 1) {
      const int z = tile / p.m_blocks, mb = tile - z * p.m_blocks;
      const int row = mb * BM + row_in_tile;
      const bool row_ok = row < p.T;
      const long long roff = (static_cast<long long>(z) * p.T + row) * p.Tp;
      if (p.mode == 1 && row_ok) {
        // backward: this thread's half row of P (written by the forward pass long ago: HBM) is requested into L2 now, while
        // the tile's MMAs run — the two sweeps below then wait for L2 hits instead of three DRAM round trips each
        // (ncu: 39 % of the kernel's stall samples sat on these loads)
        const int pc0 = c_lo * 32, pc1 = min(c_hi * 32, p.Tp);
        if (pc1 > pc0)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.p_in + roff + pc0), "r"((pc1 - pc0) * 2) : "memory");
      }
      mbar_wait(tfull_bar, acc_phase);
      tc_fence_after();
      constexpr int G = 2;  // (320 threads: 168 registers each — two 32-column chunks in flight per thread)
      uint32_t rg[G][32];
      auto load_group = [&](int cb) {  // chunks cb .. cb+G-1 of this thread's range
#pragma unroll
        for (int i = 0; i < G; ++i)
          if (cb + i < c_hi) tmem_ld_32x32(t_addr + (cb + i) * 32, rg[i]);
        tmem_ld_wait();
      };
      if (p.mode == 0) {
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        for (int cb = c_lo; cb < c_hi; cb += G) {
          if (cb * 32 >= p.T) break;
          load_group(cb);
#pragma unroll
          for (int i = 0; i < G; ++i) {
            const int col0 = (cb + i) * 32;
            if (cb + i >= c_hi) continue;
            if (col0 + 32 <= p.T) {
#pragma unroll
              for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(rg[i][j]));
            } else if (col0 < p.T) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.T) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(rg[i][j]));
            }
          }
        }
        float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        m = fmaxf(m, exchange(xch_a, m));
        const float msl = m * sl2;
        // second sweep: e = 2^(sl2 * s - msl) is computed ONCE per score, summed, and written back over the score in TMEM
        // (tcgen05.st; each thread re-reads only what it wrote) — the third sweep is then a multiply, not a second exp
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        for (int cb = c_lo; cb < c_hi; cb += G) {
          if (cb * 32 >= p.T) break;
          load_group(cb);
#pragma unroll
          for (int i = 0; i < G; ++i) {
            const int col0 = (cb + i) * 32;
            if (cb + i >= c_hi || col0 >= p.T) continue;  // (columns >= T hold exact zeros: their key rows were zero-filled)
            if (col0 + 32 <= p.T) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float e = ex2_fast(fmaf(__uint_as_float(rg[i][j]), sl2, -msl));
                s4[j & 3] += e;
                rg[i][j] = __float_as_uint(e);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float e = (col0 + j < p.T) ? ex2_fast(fmaf(__uint_as_float(rg[i][j]), sl2, -msl)) : 0.f;
                s4[j & 3] += e;
                rg[i][j] = __float_as_uint(e);
              }
            }
            tmem_st_32x32(t_addr + (cb + i) * 32, rg[i]);
          }
        }
        tmem_st_wait();
        const float mine = (s4[0] + s4[1]) + (s4[2] + s4[3]);
        const float other = exchange(xch_b, mine);
        const float inv = 1.0f / (half ? other + mine : mine + other);  // the same operand order in both threads of a row
        for (int cb = c_lo; cb < c_hi; cb += G) {
          if (cb * 32 >= p.Tp) break;
          load_group(cb);
#pragma unroll
          for (int i = 0; i < G; ++i) {
            const int col0 = (cb + i) * 32;
            if (col0 < p.Tp && cb + i < c_hi) {
              float v[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rg[i][j]) * inv;
              if (p.tma_store) {
                store_chunk(v, col0, mb * BM + q * 32, z);
              } else if (row_ok) {
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8)
                  if (col0 + g8 * 8 < p.Tp) *reinterpret_cast<uint4*>(p.out + roff + col0 + g8 * 8) = attn_pack8(v + g8 * 8);
              }
            }
          }
        }
      } else {
        const __nv_bfloat16* pr = p.p_in + roff;
        uint4 pu[G][4];
        auto load_p = [&](int cb) {
#pragma unroll
          for (int i = 0; i < G; ++i)
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              const int col = (cb + i) * 32 + g8 * 8;
              pu[i][g8] = (row_ok && cb + i < c_hi && col < p.Tp) ? __ldg(reinterpret_cast<const uint4*>(pr + col)) : make_uint4(0u, 0u, 0u, 0u);
            }
        };
        float d4[4] = {0.f, 0.f, 0.f, 0.f};
        for (int cb = c_lo; cb < c_hi; cb += G) {
          if (cb * 32 >= p.T) break;
          load_p(cb);  // global loads and TMEM loads of the group are in flight together
          load_group(cb);
#pragma unroll
          for (int i = 0; i < G; ++i) {
            if (cb + i < c_hi) {
#pragma unroll
              for (int g8 = 0; g8 < 4; ++g8) {  // padding columns of P are zero: they add nothing
                const float2 a = unpack_bf16x2(pu[i][g8].x), b2 = unpack_bf16x2(pu[i][g8].y), c2 = unpack_bf16x2(pu[i][g8].z),
                             e2 = unpack_bf16x2(pu[i][g8].w);
                const int o = g8 * 8;
                d4[0] = fmaf(a.x, __uint_as_float(rg[i][o]), d4[0]), d4[1] = fmaf(a.y, __uint_as_float(rg[i][o + 1]), d4[1]);
                d4[2] = fmaf(b2.x, __uint_as_float(rg[i][o + 2]), d4[2]), d4[3] = fmaf(b2.y, __uint_as_float(rg[i][o + 3]), d4[3]);
                d4[0] = fmaf(c2.x, __uint_as_float(rg[i][o + 4]), d4[0]), d4[1] = fmaf(c2.y, __uint_as_float(rg[i][o + 5]), d4[1]);
                d4[2] = fmaf(e2.x, __uint_as_float(rg[i][o + 6]), d4[2]), d4[3] = fmaf(e2.y, __uint_as_float(rg[i][o + 7]), d4[3]);
              }
            }
          }
        }
        const float mine = (d4[0] + d4[1]) + (d4[2] + d4[3]);
        const float other = exchange(tile_parity ? xch_b : xch_a, mine);  // one exchange per tile: alternate the area
        const float dot = half ? other + mine : mine + other;
        for (int cb = c_lo; cb < c_hi; cb += G) {
          if (cb * 32 >= p.Tp) break;
          load_p(cb);  // second sweep: the row is L1/L2-resident now
          load_group(cb);
          if (p.tma_store) {
#pragma unroll
            for (int i = 0; i < G; ++i) {
              const int col0 = (cb + i) * 32;
              if (col0 < p.Tp && cb + i < c_hi) {  // warp-uniform
                float v[32];
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {  // rows >= T / columns >= Tp were loaded as zeros (and are clipped anyway)
                  const float2 a = unpack_bf16x2(pu[i][g8].x), b2 = unpack_bf16x2(pu[i][g8].y), c2 = unpack_bf16x2(pu[i][g8].z),
                               e2 = unpack_bf16x2(pu[i][g8].w);
                  const float pv[8] = {a.x, a.y, b2.x, b2.y, c2.x, c2.y, e2.x, e2.y};
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[g8 * 8 + j] = pv[j] * (__uint_as_float(rg[i][g8 * 8 + j]) - dot) * p.scale;
                }
                store_chunk(v, col0, mb * BM + q * 32, z);
              }
            }
          } else if (row_ok) {
#pragma unroll
            for (int i = 0; i < G; ++i) {
#pragma unroll
              for (int g8 = 0; g8 < 4; ++g8) {
                const int col = (cb + i) * 32 + g8 * 8;
                if (col < p.Tp && cb + i < c_hi) {
                  const float2 a = unpack_bf16x2(pu[i][g8].x), b2 = unpack_bf16x2(pu[i][g8].y), c2 = unpack_bf16x2(pu[i][g8].z),
                               e2 = unpack_bf16x2(pu[i][g8].w);
                  const float pv[8] = {a.x, a.y, b2.x, b2.y, c2.x, c2.y, e2.x, e2.y};
                  float v[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) v[j] = pv[j] * (__uint_as_float(rg[i][g8 * 8 + j]) - dot) * p.scale;
                  *reinterpret_cast<uint4*>(p.out + roff + col) = attn_pack8(v);
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar);
      acc_phase ^= 1;
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

// ------------------------------------------------------------------------------------------------ fused forward
// Flash-style forward of one (batch, head, 128-query block) per tile: S = Q K^T in TMEM -> row softmax by the epilogue
// threads (as in mode 0) -> P written as bf16 into SHARED memory in the canonical K-major SWIZZLE_128B layout -> the MMA
// thread multiplies it with V (streamed through the same TMA ring as MN-major B tiles) into TMEM columns that re-use
// the score columns -> the epilogue drains O.  P reaches HBM only when the caller wants it (training: the backward
// reads it); an evaluation pass never writes it.  One launch instead of two (attn_scores + the batched P.V GEMM), Q is
// loaded once, P never round-trips.  Requires 160 < T <= 320 (two score parts, five 64-key blocks of P) and a head
// dim that is 64, 128, 192, 256 or 2 x {64..256 step 64} (O is accumulated as one or two N-halves).
constexpr int kFwdStages = 2;
constexpr int kFwdPBytes = 5 * BM * BK * 2;  // 80 KiB: P tile 128 x 320 bf16 as five K-major SW128 k-blocks
constexpr int kFwdSmem = kFwdStages * kAttnStage + kFwdPBytes + 256 + 1024;

struct alignas(64) AttnFwdParams {
  CUtensorMap tmq, tmk, tmv;
  int T, Tp, heads, dh, num_kb, m_blocks, num_tiles;
  int q_off, k_off, v_off;
  int n_halves, half_n;
  float scale;
  __nv_bfloat16* p_out;  // optional (Z, T, Tp)
  __nv_bfloat16* o_out;  // (n_batch * T, o_ld), head h at columns o_off + h * dh
  long long o_ld;
  int o_off;
  uint32_t k_lbo, k_sbo, mn_lbo, mn_sbo;
};

__global__ void __launch_bounds__(kGemmThreads, 1) attn_fwd_kernel(const __grid_constant__ AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* p_s = smem + kFwdStages * kAttnStage;  // 1 KiB aligned (stage size is a multiple of 1 KiB)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(p_s + kFwdPBytes);
  uint64_t* empty_bar = full_bar + kFwdStages;
  uint64_t* sfull_bar = empty_bar + kFwdStages;  // MMA -> epilogue: scores complete
  uint64_t* pready_bar = sfull_bar + 1;          // epilogue (128 threads) -> MMA: P tile in shared memory, S columns free
  uint64_t* ofull_bar = pready_bar + 1;          // MMA -> epilogue: O complete
  uint64_t* tempty_bar = ofull_bar + 1;          // epilogue (128 threads) -> MMA: TMEM free for the next tile
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tempty_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  asm volatile("griddepcontrol.launch_dependents;");
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tmq);
    prefetch_tmap(&p.tmk);
    prefetch_tmap(&p.tmv);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kFwdStages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(sfull_bar, 1);
      mbar_init(pready_bar, 128);
      mbar_init(ofull_bar, 1);
      mbar_init(tempty_bar, 128);
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
  constexpr int kVkb = 5;  // 64-key blocks of the P.V contraction (keys >= T: zero columns of P, zero-filled rows of V)
  const uint32_t v_tx = static_cast<uint32_t>(p.dh) * BK * 2;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int z = tile / p.m_blocks, mb = tile - z * p.m_blocks;
        const int b = z / p.heads, h = z - b * p.heads;
        for (int kb = 0; kb < p.num_kb; ++kb) {  // Q and K over the head dims
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kAttnStage;
          uint8_t* sb = sa + kAttnAB;
          mbar_expect_tx(&full_bar[stage], kAttnAB + 2 * kAttnBB);
          tma_load_3d(sa, &p.tmq, &full_bar[stage], p.q_off + h * p.dh + kb * BK, mb * BM, b);
          for (int j = 0; j < 2; ++j) tma_load_3d(sb + j * kAttnBB, &p.tmk, &full_bar[stage], p.k_off + h * p.dh + kb * BK, j * kAttnNP, b);
          if (++stage == kFwdStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        for (int kb = 0; kb < kVkb; ++kb) {  // V over the keys: dh / 64 boxes of (64 dims x 64 keys), MN-major B tiles
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sv = smem + stage * kAttnStage;
          mbar_expect_tx(&full_bar[stage], v_tx);
          for (int j = 0; j < p.dh / 64; ++j) tma_load_3d(sv + j * (BK * 128), &p.tmv, &full_bar[stage], p.v_off + h * p.dh + j * 64, kb * BK, b);
          if (++stage == kFwdStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(BM, kAttnNP, false, false);
      const uint32_t idesc_o = make_idesc_bf16(BM, p.half_n, false, true);
      int stage = 0;
      uint32_t phase = 0, tphase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar, tphase ^ 1);
        tc_fence_after();
        uint32_t accumulate = 0;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kAttnStage);
          const uint32_t sb = sa + kAttnAB;
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t da = make_smem_desc(sa + kk * 32, p.k_lbo, p.k_sbo);
            for (int j = 0; j < 2; ++j) {
              const uint64_t db = make_smem_desc(sb + j * kAttnBB + kk * 32, p.k_lbo, p.k_sbo);
              umma_bf16(tmem_base + j * kAttnNP, da, db, idesc_s, accumulate);
            }
            accumulate = 1;
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == kFwdStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(sfull_bar);
        // ---- O = P V: A = the P tile the epilogue threads wrote, B = V k-blocks from the ring, D re-uses the S columns
        mbar_wait(pready_bar, tphase);
        tc_fence_after();
        accumulate = 0;
        for (int kb = 0; kb < kVkb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sp = smem_u32(p_s + kb * (BM * BK * 2));
          const uint32_t sv = smem_u32(smem + stage * kAttnStage);
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk) {
            const uint64_t da = make_smem_desc(sp + kk * 32, p.k_lbo, p.k_sbo);
            for (int hf = 0; hf < p.n_halves; ++hf) {
              const uint64_t db = make_smem_desc(sv + hf * (p.half_n / 64) * (BK * 128) + kk * (16 * 128), p.mn_lbo, p.mn_sbo);
              umma_bf16(tmem_base + hf * p.half_n, da, db, idesc_o, accumulate);
            }
            accumulate = 1;
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == kFwdStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(ofull_bar);
        tphase ^= 1;
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int row_in_tile = q * 32 + lane;
    const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    constexpr int nchunks = 2 * (kAttnNP / 32);  // 10 chunks = 320 score columns
    const float sl2 = p.scale * 1.4426950408889634f;
    uint8_t* p_row = p_s + row_in_tile * 128;    // this query row inside every k-block tile
    const int swz = row_in_tile & 7;
    uint32_t tphase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int z = tile / p.m_blocks, mb = tile - z * p.m_blocks;
      const int b = z / p.heads, h = z - b * p.heads;
      const int row = mb * BM + row_in_tile;
      const bool row_ok = row < p.T;
      const long long roff = (static_cast<long long>(z) * p.T + row) * p.Tp;
      mbar_wait(sfull_bar, tphase);
      tc_fence_after();
      constexpr int G = 4;
      uint32_t rg[G][32];
      constexpr int ngroups = (nchunks + G - 1) / G;
      auto load_group = [&](int g) {
#pragma unroll
        for (int i = 0; i < G; ++i)
          if (g * G + i < nchunks) tmem_ld_32x32(t_addr + (g * G + i) * 32, rg[i]);
        tmem_ld_wait();
      };
      float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      for (int g = 0; g < ngroups; ++g) {
        if (g * G * 32 >= p.T) break;
        load_group(g);
#pragma unroll
        for (int i = 0; i < G; ++i) {
          const int col0 = (g * G + i) * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.T) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(rg[i][j]));
        }
      }
      const float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      const float msl = m * sl2;
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
      for (int g = 0; g < ngroups; ++g) {
        if (g * G * 32 >= p.T) break;
        load_group(g);
#pragma unroll
        for (int i = 0; i < G; ++i) {
          const int col0 = (g * G + i) * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + j < p.T) s4[j & 3] += ex2_fast(fmaf(__uint_as_float(rg[i][j]), sl2, -msl));
        }
      }
      const float inv = 1.0f / ((s4[0] + s4[1]) + (s4[2] + s4[3]));
      for (int g = 0; g < ngroups; ++g) {  // every one of the 320 columns goes to shared memory (zeros past T)
        load_group(g);
#pragma unroll
        for (int i = 0; i < G; ++i) {
          if (g * G + i >= nchunks) continue;
          const int col0 = (g * G + i) * 32;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = (col0 + j < p.T) ? ex2_fast(fmaf(__uint_as_float(rg[i][j]), sl2, -msl)) * inv : 0.f;
          uint8_t* kb_row = p_row + (col0 >> 6) * (BM * BK * 2);
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            const uint4 pk = attn_pack8(v + g8 * 8);
            const int c = ((col0 & 63) >> 3) + g8;  // 16-byte chunk inside the 128-byte row of this k-block
            *reinterpret_cast<uint4*>(kb_row + ((c ^ swz) << 4)) = pk;
            if (p.p_out && row_ok && col0 + g8 * 8 < p.Tp) *reinterpret_cast<uint4*>(p.p_out + roff + col0 + g8 * 8) = pk;
          }
        }
      }
      fence_proxy_async();  // the P tile was written through the generic proxy; tcgen05.mma reads it through the async proxy
      tc_fence_before();
      mbar_arrive(pready_bar);
      // ---- drain O
      mbar_wait(ofull_bar, tphase);
      tc_fence_after();
      __nv_bfloat16* orow = p.o_out + (static_cast<long long>(b) * p.T + row) * p.o_ld + p.o_off + h * p.dh;
      const int ochunks = p.dh / 32;
      for (int g = 0; g * G < ochunks; ++g) {
#pragma unroll
        for (int i = 0; i < G; ++i)
          if (g * G + i < ochunks) tmem_ld_32x32(t_addr + (g * G + i) * 32, rg[i]);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < G; ++i) {
            if (g * G + i >= ochunks) continue;
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rg[i][j]);
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) *reinterpret_cast<uint4*>(orow + (g * G + i) * 32 + g8 * 8) = attn_pack8(v + g8 * 8);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar);
      tphase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace tribe

extern "C" int tribe_attn_scores(const void* a, int64_t a_ld, int64_t a_off, const void* b, int64_t b_ld, int64_t b_off, int64_t n_batch,
                                 int64_t T, int64_t heads, int64_t dh, float scale, int32_t mode, const void* p_in, void* out, int64_t Tp,
                                 void* stream) {
  using namespace tribe;
  if (!a || !b || !out || n_batch <= 0 || T <= 0 || heads <= 0 || dh <= 0 || (mode != 0 && mode != 1) || (mode == 1 && !p_in))
    return set_error(TRIBE_EINVAL, "attn_scores: bad arguments");
  if (dh % BK || T > 2 * kAttnNP || Tp < T || Tp > 2 * kAttnNP || Tp % 8)
    return set_error(TRIBE_EINVAL, "attn_scores: needs head_dim % 64 == 0, T <= Tp <= 320, Tp % 8 == 0");
  if ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(p_in)) & 15) return set_error(TRIBE_EINVAL, "attn_scores: P / out must be 16-byte aligned");
  AttnKParams kp;
  memset(&kp, 0, sizeof(kp));
  TribeOperand oa, ob;
  memset(&oa, 0, sizeof(oa));
  memset(&ob, 0, sizeof(ob));
  oa.ptr = a, oa.inner = a_ld, oa.rows = T, oa.batch = n_batch, oa.row_stride = a_ld, oa.batch_stride = T * a_ld;
  ob.ptr = b, ob.inner = b_ld, ob.rows = T, ob.batch = n_batch, ob.row_stride = b_ld, ob.batch_stride = T * b_ld;
  int rc = encode_operand(oa, BM, &kp.tma);
  if (rc) return rc;
  rc = encode_operand(ob, kAttnNP, &kp.tmb);
  if (rc) return rc;
  kp.T = static_cast<int>(T), kp.Tp = static_cast<int>(Tp), kp.heads = static_cast<int>(heads), kp.dh = static_cast<int>(dh);
  kp.n_parts = T > kAttnNP ? 2 : 1;
  kp.num_kb = static_cast<int>(dh / BK);
  kp.m_blocks = static_cast<int>((T + BM - 1) / BM);
  kp.num_tiles = static_cast<int>(kp.m_blocks * n_batch * heads);
  kp.a_off = static_cast<int>(a_off), kp.b_off = static_cast<int>(b_off);
  kp.scale = scale, kp.mode = mode;
  kp.p_in = reinterpret_cast<const __nv_bfloat16*>(p_in);
  kp.out = reinterpret_cast<__nv_bfloat16*>(out);
  kp.k_lbo = 16, kp.k_sbo = 1024;
  static const int allow_tma_store = [] {  // TRIBE_TMA_STORE=0: row-per-thread 16-byte stores (the round-1 epilogue)
    const char* e = getenv("TRIBE_TMA_STORE");
    return e ? atoi(e) : 1;
  }();
  if (allow_tma_store) {
    rc = encode_out_bf16_3d(out, Tp, T, n_batch * heads, Tp, T * Tp, &kp.tmo);
    if (rc) return rc;
    kp.tma_store = 1;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(attn_scores)");
    attr_set = true;
  }
  const int grid = kp.num_tiles < num_sms() ? kp.num_tiles : num_sms();
  {
    cudaError_t le = launch_k(attn_scores_kernel, dim3(grid), dim3(kAttnThreads), kAttnSmem, reinterpret_cast<cudaStream_t>(stream), kp);
    if (le != cudaSuccess) return set_cuda_error(le, "attn_scores launch");
  }
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e, "attn_scores launch");
  return TRIBE_OK;
}

extern "C" int tribe_attn_fwd(const void* q, int64_t q_ld, int64_t q_off, const void* k, int64_t k_ld, int64_t k_off, const void* v, int64_t v_ld,
                              int64_t v_off, int64_t n_batch, int64_t T, int64_t heads, int64_t dh, float scale, void* p_out, int64_t Tp,
                              void* o_out, int64_t o_ld, int64_t o_off, void* stream) {
  using namespace tribe;
  if (!q || !k || !v || !o_out || n_batch <= 0 || T <= 0 || heads <= 0 || dh <= 0) return set_error(TRIBE_EINVAL, "attn_fwd: bad arguments");
  if (dh % BK || T <= kAttnNP || T > 2 * kAttnNP || Tp < T || Tp > 2 * kAttnNP || Tp % 8)
    return set_error(TRIBE_EINVAL, "attn_fwd: needs head_dim % 64 == 0, 160 < T <= Tp <= 320, Tp % 8 == 0");
  const int n_halves = dh > 256 ? 2 : 1;
  const int64_t half_n = dh / n_halves;
  if (half_n % 64 || half_n > 256 || dh * BK * 2 > kAttnStage) return set_error(TRIBE_EINVAL, "attn_fwd: head_dim must be 64..256 or 2 x (64..256), multiples of 64");
  if ((reinterpret_cast<uintptr_t>(o_out) | reinterpret_cast<uintptr_t>(p_out)) & 15 || o_ld % 8 || o_off % 8)
    return set_error(TRIBE_EINVAL, "attn_fwd: outputs must be 16-byte aligned");
  AttnFwdParams kp;
  memset(&kp, 0, sizeof(kp));
  TribeOperand oq, ok, ov;
  memset(&oq, 0, sizeof(oq)), memset(&ok, 0, sizeof(ok)), memset(&ov, 0, sizeof(ov));
  oq.ptr = q, oq.inner = q_ld, oq.rows = T, oq.batch = n_batch, oq.row_stride = q_ld, oq.batch_stride = T * q_ld;
  ok.ptr = k, ok.inner = k_ld, ok.rows = T, ok.batch = n_batch, ok.row_stride = k_ld, ok.batch_stride = T * k_ld;
  ov.ptr = v, ov.inner = v_ld, ov.rows = T, ov.batch = n_batch, ov.row_stride = v_ld, ov.batch_stride = T * v_ld;
  int rc = encode_operand(oq, BM, &kp.tmq);
  if (rc) return rc;
  rc = encode_operand(ok, kAttnNP, &kp.tmk);
  if (rc) return rc;
  rc = encode_operand(ov, BK, &kp.tmv);  // MN-major operand: boxes of 64 dims x 64 keys
  if (rc) return rc;
  kp.T = static_cast<int>(T), kp.Tp = static_cast<int>(Tp), kp.heads = static_cast<int>(heads), kp.dh = static_cast<int>(dh);
  kp.num_kb = static_cast<int>(dh / BK);
  kp.m_blocks = static_cast<int>((T + BM - 1) / BM);
  kp.num_tiles = static_cast<int>(kp.m_blocks * n_batch * heads);
  kp.q_off = static_cast<int>(q_off), kp.k_off = static_cast<int>(k_off), kp.v_off = static_cast<int>(v_off);
  kp.n_halves = n_halves, kp.half_n = static_cast<int>(half_n);
  kp.scale = scale;
  kp.p_out = reinterpret_cast<__nv_bfloat16*>(p_out);
  kp.o_out = reinterpret_cast<__nv_bfloat16*>(o_out);
  kp.o_ld = o_ld, kp.o_off = static_cast<int>(o_off);
  kp.k_lbo = 16, kp.k_sbo = 1024, kp.mn_lbo = BK * 128, kp.mn_sbo = 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem);
    if (e != cudaSuccess) return set_cuda_error(e, "cudaFuncSetAttribute(attn_fwd)");
    attr_set = true;
  }
  const int grid = kp.num_tiles < num_sms() ? kp.num_tiles : num_sms();
  cudaError_t le = launch_k(attn_fwd_kernel, dim3(grid), dim3(kGemmThreads), kFwdSmem, reinterpret_cast<cudaStream_t>(stream), kp);
  count_launch();
  if (le == cudaSuccess) le = cudaGetLastError();
  if (le != cudaSuccess) return set_cuda_error(le, "attn_fwd launch");
  return TRIBE_OK;
}
