// Data-parallel step tail in ONE kernel per owned range (sm_100a + NVLink 5 / NVSwitch):
//
//   gradient reduction over ranks  ->  fused Adam on the slice of the bucket THIS rank owns  ->  multicast of the updated
//   bf16 shadow weights (multimem.st: one store lands in every rank's copy).
//
// The reduction reads either (a) `world` staged copies — every rank's copy engines PUSH its slice of a finished bucket
// into the owner's staging buffer over NVLink while the backward pass continues (tribe_memcpy_async: no SM involved), the
// kernel sums them in rank order from local HBM; (b) the NVLS multicast mapping (multimem.ld_reduce, in-switch sum); or
// (c) the peers' buffers through peer-mapped pointers.
//
// It replaces the reference's DDP gradient all-reduce + replicated optimizer step (Lightning DDP,
// algonauts2025/main.py:388-394; Adam recipe algonauts2025/grids/defaults.py:126-141): per step a rank receives 4 B per
// OWNED parameter from every peer (instead of all-reducing 4 B per parameter both ways), touches the optimizer state of
// 1/N of the model, and receives 2 B per parameter of shadow weights.
//
// Measured on 2 x B200 (profiles/r02_xgpu_probe*.log, r02_coresidency.log): multimem.ld_reduce.v4.f32 tops out at
// ~166 GB/s of reduced data per GPU and multimem.st at ~340 GB/s whatever the parallelism; peer loads reach ~330 GB/s;
// and ANY foreign CTA that stays resident on an SM (even 32 threads) keeps the persistent 2-CTA tcgen05 GEMM off that
// SM, whose static tile schedule then pays a second wave (1.5x per GEMM).  Hence the default: gradients travel by copy
// engine during the backward pass, and this kernel runs after it, at HBM speed, with the device to itself.
//
// Cross-rank ordering uses xgpu_barrier_kernel: flag words in a symmetric buffer, one slot per (bucket, source rank),
// CAS flip-flop (0 -> 1 by the signaller with release.sys, 1 -> 0 by the waiter with acquire.sys), so no epoch numbers
// are needed and a captured CUDA graph can replay it.  Every spin has a wall-clock timeout (globaltimer) that raises a
// sticky error word instead of hanging the GPU.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "adam_math.cuh"
#include "tribe_b200.h"
#include "tribe_internal.h"

namespace tribe {

__device__ __forceinline__ uint64_t xg_now() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ uint32_t xg_cas_release(uint32_t* addr, uint32_t expect, uint32_t set) {
  uint32_t old;
  asm volatile("atom.release.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(expect), "r"(set) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t xg_cas_acquire(uint32_t* addr, uint32_t expect, uint32_t set) {
  uint32_t old;
  asm volatile("atom.acquire.sys.global.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(expect), "r"(set) : "memory");
  return old;
}

// thread j: tell rank j "rank `rank` has arrived at `slot`", then wait for rank j's arrival.  All device-scope writes of
// earlier kernels of this stream happen-before the release (kernel boundary + fence), so after the barrier every rank
// may read what the others wrote before it (gradients) and overwrite what the others read before it (shadow weights).
__global__ void __launch_bounds__(32) xgpu_barrier_kernel(TribeXgpuPeers flags, int rank, int world, int slot, uint32_t* err, uint64_t timeout_ns) {
  const int j = threadIdx.x;
  __threadfence_system();
  if (j < world && j != rank) {
    uint32_t* remote = reinterpret_cast<uint32_t*>(flags.ptr[j]) + slot * TRIBE_XGPU_MAX_WORLD + rank;
    uint32_t* local = reinterpret_cast<uint32_t*>(flags.ptr[rank]) + slot * TRIBE_XGPU_MAX_WORLD + j;
    const uint64_t t0 = xg_now();
    bool ok = true;
    while (xg_cas_release(remote, 0u, 1u) != 0u) {
      if (xg_now() - t0 > timeout_ns) {
        ok = false;
        break;
      }
    }
    while (ok && xg_cas_acquire(local, 1u, 0u) != 1u) {
      if (xg_now() - t0 > timeout_ns) {
        ok = false;
        break;
      }
    }
    if (!ok) atomicExch(err, 1u + static_cast<uint32_t>(slot));
  }
  __threadfence_system();
}

struct ShardedAdamK {
  float* p;
  float* m;
  float* v;
  const float* g_mc;       // multicast address of the owned gradient range (MC)
  uint16_t* s_mc;          // multicast address of the owned shadow range (MC)
  float* p_mc;             // multicast address of the owned master range (MC, bcast only)
  TribeXgpuPeers g_peer;   // per-rank addresses of the same ranges (P2P)
  TribeXgpuPeers s_peer;
  TribeXgpuPeers p_peer;
  int64_t n;               // elements, multiple of 8
  int world, rank, bcast;
  float inv_world;
  const float* hyper;      // {beta1, beta2, lr / bc1, 1 / sqrt(bc2), eps, weight_decay} in device memory
};

__device__ __forceinline__ float4 xg_ld_reduce(const float* mc) {
  float4 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(mc)
               : "memory");
  return r;
}
__device__ __forceinline__ void xg_mc_store(void* mc, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" : : "l"(mc), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 xg_ld_peer(const float* p) {
  float4 r;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ void xg_st_peer(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.relaxed.sys.global.v4.b32 [%0], {%1, %2, %3, %4};" : : "l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ uint32_t xg_pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// identical arithmetic to adam_kernel (optim.cu) — the shared adam_math.cuh: the single-GPU and the sharded step produce
// the same bits from the same gradient
__device__ __forceinline__ void xg_adam_one(float& p, float g, float& m, float& v, float beta1, float beta2, float step_size, float inv_bc2_sqrt,
                                            float eps, float wd) {
  adam_one(p, g, m, v, beta1, beta2, step_size, inv_bc2_sqrt, eps, wd);
}

// GMC: gradients through multimem.ld_reduce (in-switch sum) vs a table of `world` addresses summed in rank order (peer
// mappings, or LOCAL staging buffers the peers' copy engines filled during the backward pass).  OMC: outputs through
// multimem.st vs per-peer stores.
template <bool GMC, bool OMC>
__global__ void __launch_bounds__(128, 6) sharded_adam_kernel(const ShardedAdamK a) {
  const float beta1 = a.hyper[0], beta2 = a.hyper[1], step_size = a.hyper[2], inv_bc2_sqrt = a.hyper[3], eps = a.hyper[4], wd = a.hyper[5];
  const int64_t nv = a.n >> 3;  // groups of 8 parameters: 2 x 16 B of gradient, one 16 B shadow store
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nv; i += stride) {
    const int64_t e = i << 3;
    float4 g0, g1;
    if (GMC) {
      g0 = xg_ld_reduce(a.g_mc + e);
      g1 = xg_ld_reduce(a.g_mc + e + 4);
    } else {
      g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0;
      for (int r = 0; r < a.world; ++r) {  // fixed rank order: every owner sums its slice the same way on every step
        const float* gp = reinterpret_cast<const float*>(a.g_peer.ptr[r]) + e;
        const float4 x0 = xg_ld_peer(gp), x1 = xg_ld_peer(gp + 4);
        g0.x += x0.x, g0.y += x0.y, g0.z += x0.z, g0.w += x0.w;
        g1.x += x1.x, g1.y += x1.y, g1.z += x1.z, g1.w += x1.w;
      }
    }
    float4 p0 = __ldcs(reinterpret_cast<const float4*>(a.p + e)), p1 = __ldcs(reinterpret_cast<const float4*>(a.p + e + 4));
    float4 m0 = __ldcs(reinterpret_cast<const float4*>(a.m + e)), m1 = __ldcs(reinterpret_cast<const float4*>(a.m + e + 4));
    float4 v0 = __ldcs(reinterpret_cast<const float4*>(a.v + e)), v1 = __ldcs(reinterpret_cast<const float4*>(a.v + e + 4));
    const float s = a.inv_world;  // gradient MEAN over ranks (DDP semantics); exact for power-of-two worlds
    xg_adam_one(p0.x, g0.x * s, m0.x, v0.x, beta1, beta2, step_size, inv_bc2_sqrt, eps, wd);
    xg_adam_one(p0.y, g0.y * s, m0.y, v0.y, beta1, beta2, step_size, inv_bc2_sqrt, eps, wd);
    xg_adam_one(p0.z, g0.z * s, m0.z, v0.z, beta1, beta2, step_size, inv_bc2_sqrt, eps, wd);
    xg_adam_one(p0.w, g0.w * s, m0.w, v0.w, beta1, beta2, step_size, inv_bc2_sqrt, eps, wd);
    xg_adam_one(p1.x, g1.x * s, m1.x, v1.x, beta1, beta2, step_size, inv_bc2_sqrt, eps, wd);
    xg_adam_one(p1.y, g1.y * s, m1.y, v1.y, beta1, beta2, step_size, inv_bc2_sqrt, eps, wd);
    xg_adam_one(p1.z, g1.z * s, m1.z, v1.z, beta1, beta2, step_size, inv_bc2_sqrt, eps, wd);
    xg_adam_one(p1.w, g1.w * s, m1.w, v1.w, beta1, beta2, step_size, inv_bc2_sqrt, eps, wd);
    __stcs(reinterpret_cast<float4*>(a.m + e), m0), __stcs(reinterpret_cast<float4*>(a.m + e + 4), m1);
    __stcs(reinterpret_cast<float4*>(a.v + e), v0), __stcs(reinterpret_cast<float4*>(a.v + e + 4), v1);
    const uint32_t s0 = xg_pack2(p0.x, p0.y), s1 = xg_pack2(p0.z, p0.w), s2 = xg_pack2(p1.x, p1.y), s3 = xg_pack2(p1.z, p1.w);
    if (OMC) {
      xg_mc_store(a.s_mc + e, s0, s1, s2, s3);
      if (a.bcast) {  // parameters the kernels read as fp32 (biases, norm gains, residual scales, positional embedding)
        xg_mc_store(a.p_mc + e, __float_as_uint(p0.x), __float_as_uint(p0.y), __float_as_uint(p0.z), __float_as_uint(p0.w));
        xg_mc_store(a.p_mc + e + 4, __float_as_uint(p1.x), __float_as_uint(p1.y), __float_as_uint(p1.z), __float_as_uint(p1.w));
      } else {
        __stcs(reinterpret_cast<float4*>(a.p + e), p0), __stcs(reinterpret_cast<float4*>(a.p + e + 4), p1);
      }
    } else {
      for (int r = 0; r < a.world; ++r) {
        xg_st_peer(reinterpret_cast<uint16_t*>(a.s_peer.ptr[r]) + e, s0, s1, s2, s3);
        if (a.bcast && r != a.rank) {
          float* pp = reinterpret_cast<float*>(a.p_peer.ptr[r]) + e;
          xg_st_peer(pp, __float_as_uint(p0.x), __float_as_uint(p0.y), __float_as_uint(p0.z), __float_as_uint(p0.w));
          xg_st_peer(pp + 4, __float_as_uint(p1.x), __float_as_uint(p1.y), __float_as_uint(p1.z), __float_as_uint(p1.w));
        }
      }
      __stcs(reinterpret_cast<float4*>(a.p + e), p0), __stcs(reinterpret_cast<float4*>(a.p + e + 4), p1);
    }
  }
}

// Micro-benchmark of the pieces (tools/xgpu_probe.py): mode 0 = multimem.ld_reduce only, 1 = peer loads only (world
// peers), 2 = local p/m/v stream only (6 x 16 B loads + 6 stores), 3 = multimem.st bf16 only, 4 = ld_reduce with UNROLL 4.
__global__ void __launch_bounds__(128, 6) xgpu_probe_kernel(const ShardedAdamK a, int mode, float* sink) {
  const int64_t nv = a.n >> 3;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  float acc = 0.f;
  if (mode == 4) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nv; i += 4 * stride) {
      float4 g[8];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t e = (i + u * stride < nv ? i + u * stride : i) << 3;
        g[2 * u] = xg_ld_reduce(a.g_mc + e), g[2 * u + 1] = xg_ld_reduce(a.g_mc + e + 4);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += g[u].x + g[u].y + g[u].z + g[u].w;
    }
  } else {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nv; i += stride) {
      const int64_t e = i << 3;
      if (mode == 0) {
        const float4 g0 = xg_ld_reduce(a.g_mc + e), g1 = xg_ld_reduce(a.g_mc + e + 4);
        acc += g0.x + g0.y + g0.z + g0.w + g1.x + g1.y + g1.z + g1.w;
      } else if (mode == 1) {
        for (int r = 0; r < a.world; ++r) {
          const float* gp = reinterpret_cast<const float*>(a.g_peer.ptr[r]) + e;
          const float4 g0 = xg_ld_peer(gp), g1 = xg_ld_peer(gp + 4);
          acc += g0.x + g0.y + g0.z + g0.w + g1.x + g1.y + g1.z + g1.w;
        }
      } else if (mode == 2) {
        float4 p0 = __ldcs(reinterpret_cast<const float4*>(a.p + e)), p1 = __ldcs(reinterpret_cast<const float4*>(a.p + e + 4));
        float4 m0 = __ldcs(reinterpret_cast<const float4*>(a.m + e)), m1 = __ldcs(reinterpret_cast<const float4*>(a.m + e + 4));
        float4 v0 = __ldcs(reinterpret_cast<const float4*>(a.v + e)), v1 = __ldcs(reinterpret_cast<const float4*>(a.v + e + 4));
        p0.x += m0.x * v0.x, p1.x += m1.x * v1.x;
        __stcs(reinterpret_cast<float4*>(a.p + e), p0), __stcs(reinterpret_cast<float4*>(a.p + e + 4), p1);
        __stcs(reinterpret_cast<float4*>(a.m + e), m0), __stcs(reinterpret_cast<float4*>(a.m + e + 4), m1);
        __stcs(reinterpret_cast<float4*>(a.v + e), v0), __stcs(reinterpret_cast<float4*>(a.v + e + 4), v1);
      } else if (mode == 3) {
        xg_mc_store(a.s_mc + e, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
      }
    }
  }
  if (acc == 123.456f) sink[0] = acc;
}

// occupies `blocks` CTAs of `threads` threads for `ns` nanoseconds (tools/coresidency_probe.py: can a small CTA share an
// SM with the persistent tcgen05 GEMM CTA?)
__global__ void xgpu_spin_kernel(uint64_t ns, uint32_t* sink) {
  const uint64_t t0 = xg_now();
  while (xg_now() - t0 < ns) {
  }
  if (ns == 1) sink[0] = 1;
}

}  // namespace tribe

using namespace tribe;

extern "C" int tribe_debug_spin(int32_t blocks, int32_t threads, double seconds, int32_t carveout_pct, uint32_t* sink, void* stream) {
  if (blocks <= 0 || threads <= 0 || threads > 1024 || !sink) return set_error(TRIBE_EINVAL, "debug_spin: bad arguments");
  if (carveout_pct >= 0) cudaFuncSetAttribute(xgpu_spin_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, carveout_pct);
  xgpu_spin_kernel<<<blocks, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(static_cast<uint64_t>(seconds * 1e9), sink);
  TRIBE_CHECK_LAUNCH("debug_spin");
  return TRIBE_OK;
}

extern "C" int tribe_xgpu_probe(const TribeShardedAdam* a, int32_t mode, int32_t blocks, float* sink, void* stream) {
  if (!a || !sink || a->n <= 0 || (a->n & 7)) return set_error(TRIBE_EINVAL, "xgpu_probe: bad arguments");
  ShardedAdamK k;
  k.p = a->param, k.m = a->m, k.v = a->v;
  k.g_mc = a->grad_mc, k.s_mc = reinterpret_cast<uint16_t*>(a->shadow_mc), k.p_mc = a->param_mc;
  k.g_peer = a->grad_peer, k.s_peer = a->shadow_peer, k.p_peer = a->param_peer;
  k.n = a->n, k.world = a->world, k.rank = a->rank, k.bcast = 0, k.inv_world = 1.f, k.hyper = a->hyper;
  xgpu_probe_kernel<<<blocks > 0 ? blocks : 148, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(k, mode, sink);
  TRIBE_CHECK_LAUNCH("xgpu_probe");
  return TRIBE_OK;
}

extern "C" int tribe_xgpu_barrier(const TribeXgpuPeers* flags, int32_t rank, int32_t world, int32_t slot, uint32_t* err_flag, double timeout_s,
                                  void* stream) {
  if (!flags || !err_flag || world < 1 || world > TRIBE_XGPU_MAX_WORLD || rank < 0 || rank >= world || slot < 0 || slot >= TRIBE_XGPU_SLOTS)
    return set_error(TRIBE_EINVAL, "xgpu_barrier: bad arguments");
  for (int r = 0; r < world; ++r)
    if (!flags->ptr[r]) return set_error(TRIBE_EINVAL, "xgpu_barrier: missing peer flag pointer");
  const uint64_t ns = static_cast<uint64_t>((timeout_s > 0 ? timeout_s : 30.0) * 1e9);
  xgpu_barrier_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*flags, rank, world, slot, err_flag, ns);
  TRIBE_CHECK_LAUNCH("xgpu_barrier");
  return TRIBE_OK;
}

extern "C" int tribe_sharded_adam_step(const TribeShardedAdam* a, void* stream) {
  if (!a || !a->param || !a->m || !a->v || !a->hyper || a->n <= 0 || (a->n & 7) || a->world < 1 || a->world > TRIBE_XGPU_MAX_WORLD ||
      a->rank < 0 || a->rank >= a->world)
    return set_error(TRIBE_EINVAL, "sharded_adam: bad arguments (n must be a positive multiple of 8)");
  const bool gmc = a->grad_mc != nullptr, omc = a->shadow_mc != nullptr;
  uintptr_t al = reinterpret_cast<uintptr_t>(a->param) | reinterpret_cast<uintptr_t>(a->m) | reinterpret_cast<uintptr_t>(a->v);
  if (gmc && !omc) return set_error(TRIBE_EINVAL, "sharded_adam: multicast gradients need multicast outputs");
  if (omc) {
    if (a->bcast_master && !a->param_mc) return set_error(TRIBE_EINVAL, "sharded_adam: multicast pointers missing");
    al |= reinterpret_cast<uintptr_t>(a->grad_mc) | reinterpret_cast<uintptr_t>(a->shadow_mc) | reinterpret_cast<uintptr_t>(a->param_mc);
  }
  for (int r = 0; r < a->world; ++r) {
    if (!gmc) {
      if (!a->grad_peer.ptr[r]) return set_error(TRIBE_EINVAL, "sharded_adam: gradient pointers missing");
      al |= reinterpret_cast<uintptr_t>(a->grad_peer.ptr[r]);
    }
    if (!omc) {
      if (!a->shadow_peer.ptr[r] || (a->bcast_master && !a->param_peer.ptr[r])) return set_error(TRIBE_EINVAL, "sharded_adam: peer pointers missing");
      al |= reinterpret_cast<uintptr_t>(a->shadow_peer.ptr[r]) | reinterpret_cast<uintptr_t>(a->param_peer.ptr[r]);
    }
  }
  if (al & 15) return set_error(TRIBE_EINVAL, "sharded_adam: every range must be 16-byte aligned");
  ShardedAdamK k;
  k.p = a->param, k.m = a->m, k.v = a->v;
  k.g_mc = a->grad_mc, k.s_mc = reinterpret_cast<uint16_t*>(a->shadow_mc), k.p_mc = a->param_mc;
  k.g_peer = a->grad_peer, k.s_peer = a->shadow_peer, k.p_peer = a->param_peer;
  k.n = a->n, k.world = a->world, k.rank = a->rank, k.bcast = a->bcast_master;
  k.inv_world = 1.0f / static_cast<float>(a->world);
  k.hyper = a->hyper;
  // the kernel runs AFTER the backward pass with the device to itself (a CTA that stays on an SM keeps the persistent
  // 2-CTA GEMM off that SM and its static tile schedule then pays a second wave: profiles/r02_coresidency.log)
  const int blocks = grid_for(a->n / 8, 128, a->max_blocks > 0 ? a->max_blocks : 148 * 6);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (gmc)
    sharded_adam_kernel<true, true><<<blocks, 128, 0, s>>>(k);
  else if (omc)
    sharded_adam_kernel<false, true><<<blocks, 128, 0, s>>>(k);
  else
    sharded_adam_kernel<false, false><<<blocks, 128, 0, s>>>(k);
  TRIBE_CHECK_LAUNCH("sharded_adam");
  return TRIBE_OK;
}

extern "C" int tribe_memcpy_async(void* dst, const void* src, int64_t n_bytes, void* stream) {
  if (!dst || !src || n_bytes < 0) return set_error(TRIBE_EINVAL, "memcpy_async: bad arguments");
  if (n_bytes == 0) return TRIBE_OK;
  // copy-engine transfer (no SM): local -> peer-mapped symmetric memory over NVLink, device-local, or device -> pinned
  // host (the direction is inferred from the unified address space)
  cudaError_t e = cudaMemcpyAsync(dst, src, static_cast<size_t>(n_bytes), cudaMemcpyDefault, reinterpret_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return set_cuda_error(e, "memcpy_async");
  return TRIBE_OK;
}
