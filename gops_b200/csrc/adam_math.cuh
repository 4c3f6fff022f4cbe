// torch.optim.Adam, single tensor path (torch/optim/adam.py _single_tensor_adam, no amsgrad / weight decay), one element.
// Every operation is pinned (no compiler-chosen contraction), so that the stand-alone Adam kernel and the gradient
// exchange kernel that applies Adam itself (peer.cu) produce bit-identical parameters.
#pragma once

namespace gops {

__device__ __forceinline__ void adam_update(float g, float& p, float& m, float& v, float omb1, float b2, float omb2, float eps,
                                            float step_size, float bc2_sqrt) {
  const float mi = fmaf(__fsub_rn(g, m), omb1, m);                          // exp_avg.lerp_(grad, 1 - beta1)
  const float vi = fmaf(__fmul_rn(g, g), omb2, __fmul_rn(v, b2));           // exp_avg_sq.mul_(beta2).addcmul_(g, g, 1 - beta2)
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vi), bc2_sqrt), eps);  // (sqrt / bias_correction2_sqrt).add_(eps)
  m = mi;
  v = vi;
  p = fmaf(-__fdiv_rn(mi, denom), step_size, p);                            // addcdiv_(exp_avg, denom, value=-step_size)
}

}  // namespace gops
