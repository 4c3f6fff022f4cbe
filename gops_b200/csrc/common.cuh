// Device-side helpers shared by the gops_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gops_b200.h"

namespace gops {

constexpr int MAXA = GOPS_B200_MAX_ACT;
constexpr int LQN = GOPS_B200_MAX_LQ_N;

// ---------------------------------------------------------------------------------------------
// mbarrier + TMA (cp.async.bulk) primitives: weights are staged global -> shared by the bulk-copy
// engine (SASS: UBLKCP) and signalled through an mbarrier transaction count.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// Activations (gops/utils/common_utils.py:26-55 -> torch.nn.{ReLU,ELU,GELU,SELU,Sigmoid,Tanh,Identity})
// h = g(x), d = g'(x) in fp32 with the accurate libdevice functions (no fast-math).
// ---------------------------------------------------------------------------------------------
// Exact-erf GELU with ONE exponential shared by the Gaussian cdf and pdf:
//   e = exp(-x^2/2);  erfc(|x|/sqrt2) = e * P(k), k = 1/(1 + p |x|/sqrt2)   (Abramowitz-Stegun 7.1.26, |err| <= 1.5e-7)
//   cdf = x >= 0 ? 1 - 0.5 e P : 0.5 e P;   pdf = e / sqrt(2 pi);   gelu = x cdf;   gelu' = cdf + x pdf
// ~20 instructions instead of libdevice erff + expf (~45); the parity tests bound the effect on loss / gradient.
__device__ __forceinline__ void gelu_parts(float x, float& cdf, float& pdf) {
  const float ax = fabsf(x);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * x * -0.72134752044448170368f));   // exp(-x^2/2)
  float k;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(k) : "f"(fmaf(ax, 0.23164189045f, 1.0f)));      // p / sqrt(2) = 0.2316419
  float pl = fmaf(1.061405429f, k, -1.453152027f);
  pl = fmaf(pl, k, 1.421413741f);
  pl = fmaf(pl, k, -0.284496736f);
  pl = fmaf(pl, k, 0.254829592f);
  const float half_erfc = 0.5f * (pl * k) * e;
  cdf = x >= 0.f ? 1.f - half_erfc : half_erfc;
  pdf = 0.39894228040143267794f * e;
}

__device__ __forceinline__ float act_fwd(int act, float x) {
  switch (act) {
    case GOPS_ACT_RELU: return fmaxf(x, 0.f);
    case GOPS_ACT_ELU: return x > 0.f ? x : expm1f(x);
    case GOPS_ACT_GELU: { float c, q; gelu_parts(x, c, q); return x * c; }
    case GOPS_ACT_SELU: return 1.0507009873554805f * (x > 0.f ? x : 1.6732632423543772f * expm1f(x));
    case GOPS_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    case GOPS_ACT_TANH: return tanhf(x);
    default: return x;
  }
}
__device__ __forceinline__ void act_fwd_grad(int act, float x, float& h, float& d) {
  switch (act) {
    case GOPS_ACT_RELU: h = fmaxf(x, 0.f); d = x > 0.f ? 1.f : 0.f; break;
    case GOPS_ACT_ELU:
      if (x > 0.f) { h = x; d = 1.f; } else { float e = expf(x); h = expm1f(x); d = e; }
      break;
    case GOPS_ACT_GELU: {
      float cdf, pdf;
      gelu_parts(x, cdf, pdf);
      h = x * cdf; d = cdf + x * pdf;
    } break;
    case GOPS_ACT_SELU: {
      const float sc = 1.0507009873554805f, al = 1.6732632423543772f;
      if (x > 0.f) { h = sc * x; d = sc; } else { h = sc * al * expm1f(x); d = sc * al * expf(x); }
    } break;
    case GOPS_ACT_SIGMOID: h = 1.f / (1.f + expf(-x)); d = h * (1.f - h); break;
    case GOPS_ACT_TANH: h = tanhf(x); d = 1.f - h * h; break;
    default: h = x; d = 1.f; break;
  }
}

// Compile-time activation variants: the per-element `switch (act)` of act_fwd / act_fwd_grad is a branch region per
// element, which stops the compiler from interleaving independent elements (every element then costs its full
// dependent-chain latency).  Hot loops dispatch ONCE on the activation id and run a loop specialised on ACT.
template <int ACT>
__device__ __forceinline__ float act_fwd_t(float x) {
  if constexpr (ACT == GOPS_ACT_RELU) return fmaxf(x, 0.f);
  else if constexpr (ACT == GOPS_ACT_ELU) return x > 0.f ? x : expm1f(x);
  else if constexpr (ACT == GOPS_ACT_GELU) { float c, q; gelu_parts(x, c, q); return x * c; }
  else if constexpr (ACT == GOPS_ACT_SELU) return 1.0507009873554805f * (x > 0.f ? x : 1.6732632423543772f * expm1f(x));
  else if constexpr (ACT == GOPS_ACT_SIGMOID) return 1.f / (1.f + expf(-x));
  else if constexpr (ACT == GOPS_ACT_TANH) return tanhf(x);
  else return x;
}
template <int ACT>
__device__ __forceinline__ void act_fwd_grad_t(float x, float& h, float& d) {
  if constexpr (ACT == GOPS_ACT_RELU) { h = fmaxf(x, 0.f); d = x > 0.f ? 1.f : 0.f; }
  else if constexpr (ACT == GOPS_ACT_ELU) {
    const float e = expf(x), em = expm1f(x);
    h = x > 0.f ? x : em; d = x > 0.f ? 1.f : e;
  } else if constexpr (ACT == GOPS_ACT_GELU) {
    float cdf, pdf;
    gelu_parts(x, cdf, pdf);
    h = x * cdf; d = cdf + x * pdf;
  } else if constexpr (ACT == GOPS_ACT_SELU) {
    const float sc = 1.0507009873554805f, al = 1.6732632423543772f;
    const float e = expf(x), em = expm1f(x);
    h = x > 0.f ? sc * x : sc * al * em; d = x > 0.f ? sc : sc * al * e;
  } else if constexpr (ACT == GOPS_ACT_SIGMOID) { h = 1.f / (1.f + expf(-x)); d = h * (1.f - h); }
  else if constexpr (ACT == GOPS_ACT_TANH) { h = tanhf(x); d = 1.f - h * h; }
  else { h = x; d = 1.f; }
}
// GOPS_ACT_SWITCH(act, M): expands M(ACT) for the runtime activation id `act` (M is a one-argument macro)
#define GOPS_ACT_SWITCH(act, M)                 \
  switch (act) {                                \
    case GOPS_ACT_RELU: M(GOPS_ACT_RELU); break;       \
    case GOPS_ACT_ELU: M(GOPS_ACT_ELU); break;         \
    case GOPS_ACT_GELU: M(GOPS_ACT_GELU); break;       \
    case GOPS_ACT_SELU: M(GOPS_ACT_SELU); break;       \
    case GOPS_ACT_SIGMOID: M(GOPS_ACT_SIGMOID); break; \
    case GOPS_ACT_TANH: M(GOPS_ACT_TANH); break;       \
    default: M(GOPS_ACT_LINEAR); break;                \
  }

// angle_normalize, gops/utils/math_utils.py:8-11: ((x + pi) % (2 pi)) - pi with floored modulo,
// evaluated in fp32 like torch.remainder on a float32 tensor.
__device__ __forceinline__ float angle_normalize(float x) {
  const float PI_F = 3.14159265358979323846f, TWO_PI_F = 6.28318530717958647692f;
  float y = __fadd_rn(x, PI_F);
  float r = fmodf(y, TWO_PI_F);
  if (r != 0.f && r < 0.f) r = __fadd_rn(r, TWO_PI_F);
  return __fsub_rn(r, PI_F);
}

}  // namespace gops
