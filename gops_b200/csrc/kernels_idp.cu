// Instantiations of the fused rollout kernel for ModelIdp (own translation unit: parallel build).
#include "kernel.cuh"

namespace gops {

typedef void (*RolloutFn)(const KParams);
typedef void (*StepFn)(const KParams, const float*, int, float*, float*, float*);

RolloutFn rollout_fn_idp(int cfg) {
  switch (cfg) {
    case 0: return rollout_kernel<ModelIdp, 128, 256>;
    case 1: return rollout_kernel<ModelIdp, 64, 256>;
    default: return rollout_kernel<ModelIdp, 32, 128>;
  }
}
StepFn step_fn_idp() { return model_step_kernel<ModelIdp>; }

}  // namespace gops
