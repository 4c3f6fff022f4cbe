// Instantiations of the fused rollout kernel for ModelLq (own translation unit: parallel build).
#include "kernel.cuh"

namespace gops {

typedef void (*RolloutFn)(const KParams);
typedef void (*StepFn)(const KParams, const float*, int, float*, float*, float*);

RolloutFn rollout_fn_lq(int cfg) {
  switch (cfg) {
    case 0: return rollout_kernel<ModelLq, 128, 256>;
    case 1: return rollout_kernel<ModelLq, 64, 256>;
    default: return rollout_kernel<ModelLq, 32, 128>;
  }
}
StepFn step_fn_lq() { return model_step_kernel<ModelLq>; }

}  // namespace gops
