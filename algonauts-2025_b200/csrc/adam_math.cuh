// One Adam update of one parameter — the arithmetic shared by the flat-range kernel (optim.cu) and the wgrad GEMM
// epilogue (gemm_common.cuh) and the rank-sharded NVLink step tail (xgpu.cu), so that all of them produce bit-identical
// parameters from bit-identical gradients.  Same formula as torch.optim.Adam
// (amsgrad=False, maximize=False), reference recipe algonauts2025/grids/defaults.py:126-141.
#pragma once
#include <cuda_runtime.h>

namespace tribe {

// hyper-parameter block in device memory: {beta1, beta2, lr / bias_correction1, 1 / sqrt(bias_correction2), eps, weight_decay, -, -}
struct AdamHyper {
  float beta1, beta2, step_size, inv_bc2_sqrt, eps, wd;
};

// Every operation is an explicitly rounded intrinsic: the compiler may not contract a multiply and an add into an FMA
// here and leave them apart there, so the three call sites cannot drift apart by an ulp (they did before: the GEMM
// epilogue and the flat kernel disagreed in the last bit of exp_avg_sq).
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float beta1, float beta2, float step_size, float inv_bc2_sqrt,
                                         float eps, float wd) {
  if (wd != 0.f) g = __fmaf_rn(wd, p, g);
  m = __fmaf_rn(__fsub_rn(g, m), __fsub_rn(1.0f, beta1), m);                             // exp_avg.lerp_(grad, 1 - beta1)
  v = __fmaf_rn(__fmul_rn(__fsub_rn(1.0f, beta2), g), g, __fmul_rn(beta2, v));           // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  const float denom = __fmaf_rn(__fsqrt_rn(v), inv_bc2_sqrt, eps);
  p = __fsub_rn(p, __fmul_rn(step_size, __fdiv_rn(m, denom)));
}

}  // namespace tribe
