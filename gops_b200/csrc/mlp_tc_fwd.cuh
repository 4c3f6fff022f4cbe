// tcgen05 / TMEM forward pass of the 64-wide policy MLP inside the fused rollout kernel (hybrid mode): the forward
// sweep's two hidden-layer GEMMs run on the 5th-generation tensor cores, the reverse sweep stays on the mma.sync
// path of rollout.cuh / mma_tiles.cuh.
//
// Status: kept as the A/B reference (GOPS_B200_ROLLOUT=hy, 9.0 ms vs 6.3 ms on C1) of the full tcgen05 kernel
// (mlp_tc_full.cuh), which runs the whole update on the tensor cores with BF16x3 operands.  This TF32 variant stops at
// the forward sweep because tcgen05 kind::tf32 reads a shared-memory operand transposed (MN-major) only in the
// SWIZZLE_128B_BASE32B layout, which no K-major read accepts: the weight-gradient products (contraction over samples)
// and the layer products (contraction over features) cannot share one fp32 operand buffer, and two copies of every
// activation of a 128-sample sub-tile do not fit in 227 KB (tools/umma_probe.cu).
//
// All NT = 512 threads cooperate on one 128-sample sub-tile: warp w owns TMEM lane quarter q = w & 3 (samples
// r = 32 q + lane) and the 16-column slice c = w >> 2 of the accumulator:
//     Xp planes <- X sub-tile;  ACC = Xp . W1^T;  epilogue -> H1 planes P;  ACC = P . W2^T;  epilogue -> output partials
// Operand planes are hi | lo pairs in the chunk-major layout of umma.cuh; the buffers alias the activation tiles of
// the mma.sync path (dead during the forward sweep).
#pragma once
#include "rollout.cuh"
#include "umma.cuh"

namespace gops {

constexpr int TC_K1 = 16;                       // layer-1 K extent (inputs padded to 16)
constexpr int TC_PLANE = 64 * 128;              // floats per hidden-activation plane
constexpr int TC_XPLANE = TC_K1 * 128;          // floats per observation plane
constexpr uint32_t TC_FWD_COLS = 64;            // TMEM columns: one [128 x 64] FP32 accumulator

struct TcCtx {
  float* W;            // staged blob: W1 hi | lo ([4][64][4] each), W2 hi | lo ([16][64][4] each), W3, b1, b2, b3
  float* Xp;           // [2][4][128][4]   observation planes
  float* P;            // [2][16][128][4]  H1 planes
  float* Zp;           // [4 slices][MAXA][128] output-layer partials
  uint64_t* bar;       // MMA completion mbarrier
  uint32_t ph;         // its parity (per thread)
  uint32_t tmem;       // TMEM base address
};

namespace tc {

struct Op {            // one MMA operand: smem byte addresses of its hi / lo planes + descriptor strides
  uint32_t hi, lo, lbo, sbo, kadv;
};
__device__ __forceinline__ Op op_k_act(const float* hi, int plane_floats) {       // activations, K-major, 128 rows
  return Op{smem_u32(hi), smem_u32(hi + plane_floats), 2048u, 128u, 4096u};
}
__device__ __forceinline__ Op op_k_w(const float* hi, int plane_floats) {         // weights, K-major, 64 rows
  return Op{smem_u32(hi), smem_u32(hi + plane_floats), 1024u, 128u, 2048u};
}

// D = A . B^T in 3xTF32 (lo.hi, hi.lo, hi.hi), KS K-steps of 8; the first MMA overwrites D
template <int KS>
__device__ __forceinline__ void issue3(uint32_t d, const Op& A, const Op& B, uint32_t idesc) {
  const uint64_t ah = umma::smem_desc(A.hi, A.lbo, A.sbo), al = umma::smem_desc(A.lo, A.lbo, A.sbo);
  const uint64_t bh = umma::smem_desc(B.hi, B.lbo, B.sbo), bl = umma::smem_desc(B.lo, B.lbo, B.sbo);
  const uint64_t ka = A.kadv >> 4, kb = B.kadv >> 4;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) umma::mma_tf32_ss(d, al + ks * ka, bh + ks * kb, idesc, ks > 0 ? 1u : 0u);
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) umma::mma_tf32_ss(d, ah + ks * ka, bl + ks * kb, idesc, 1u);
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) umma::mma_tf32_ss(d, ah + ks * ka, bh + ks * kb, idesc, 1u);
}

__device__ __forceinline__ void wait_mma(TcCtx& cx) {
  mbar_wait(cx.bar, cx.ph);
  cx.ph ^= 1u;
  umma::fence_after_sync();
}
// publish this thread's shared-memory writes / TMEM reads to the MMA issuer and meet the other threads
__device__ __forceinline__ void publish_sync() {
  fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
}

// hidden-layer epilogue of thread (q, c): 16 accumulator columns -> + bias -> activation -> hi | lo planes
template <int ACT>
__device__ __forceinline__ void epi_hidden(uint32_t t_acc, const float* __restrict__ bias16,
                                           float* __restrict__ plane_hi, int c, int r) {
  float v[16];
  umma::tmem_ld16(t_acc, v);
#pragma unroll
  for (int c4 = 0; c4 < 4; ++c4) {
    const float4 b = *reinterpret_cast<const float4*>(bias16 + 4 * c4);
    const float pb[4] = {b.x, b.y, b.z, b.w};
    float hh[4], hl[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) umma::split(act_fwd_t<ACT>(v[4 * c4 + e] + pb[e]), hh[e], hl[e]);
    reinterpret_cast<float4*>(plane_hi)[(4 * c + c4) * 128 + r] = make_float4(hh[0], hh[1], hh[2], hh[3]);
    reinterpret_cast<float4*>(plane_hi + TC_PLANE)[(4 * c + c4) * 128 + r] = make_float4(hl[0], hl[1], hl[2], hl[3]);
  }
}
// last hidden layer + this slice's share of the output dot products
template <int ACT>
__device__ __forceinline__ void epi_last(uint32_t t_acc, const float* __restrict__ bias16,
                                         const float* __restrict__ W3, int out, int c, float* zp) {
  float v[16];
  umma::tmem_ld16(t_acc, v);
#pragma unroll
  for (int e = 0; e < 16; ++e) v[e] = act_fwd_t<ACT>(v[e] + bias16[e]);
#pragma unroll
  for (int a = 0; a < MAXA; ++a) {
    zp[a] = 0.f;
    if (a < out) {
      const float* w = W3 + a * 64 + 16 * c;
#pragma unroll
      for (int e = 0; e < 16; ++e) zp[a] = fmaf(w[e], v[e], zp[a]);
    }
  }
}

}  // namespace tc

// X sub-tile ([feature][sample], leading dimension XS, `xrows` rows) -> Zout[a * XS + r], a < L.out, r < 128.
// L carries the chunk-major blob offsets (make_net_tc); ends with a CTA barrier.
template <int NT>
__device__ __forceinline__ void mlp_forward_tc(const NetL& L, TcCtx& cx, const float* __restrict__ Xsub, int XS,
                                               int xrows, float* __restrict__ Zout) {
  static_assert(NT == 512, "cooperative tcgen05 MLP: 16 warps = 4 lane quarters x 4 column slices");
  const int tid = threadIdx.x, lane = tid & 31, q = (tid >> 5) & 3, c = tid >> 7, r = 32 * q + lane;
  const uint32_t tl = cx.tmem + ((uint32_t)(32 * q) << 16) + 16 * c;      // this thread's lane quarter + column slice
  {  // observation planes: thread = (chunk tid >> 7, row tid & 127)
    const int row = tid & 127, ch = tid >> 7;
    float xh[4], xl[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int f = 4 * ch + e;
      umma::split(f < xrows ? Xsub[f * XS + row] : 0.f, xh[e], xl[e]);
    }
    reinterpret_cast<float4*>(cx.Xp)[ch * 128 + row] = make_float4(xh[0], xh[1], xh[2], xh[3]);
    reinterpret_cast<float4*>(cx.Xp + TC_XPLANE)[ch * 128 + row] = make_float4(xl[0], xl[1], xl[2], xl[3]);
  }
  tc::publish_sync();
  if (tid == 0) {
    umma::fence_after_sync();
    tc::issue3<TC_K1 / 8>(cx.tmem, tc::op_k_act(cx.Xp, TC_XPLANE), tc::op_k_w(cx.W + L.o_w1, 64 * TC_K1),
                          umma::idesc_tf32(128, 64, false, false));
    umma::commit(cx.bar);
  }
  tc::wait_mma(cx);
#define GOPS_TC_H(A) tc::epi_hidden<A>(tl, cx.W + L.o_b1 + 16 * c, cx.P, c, r)
  GOPS_ACT_SWITCH(L.hact, GOPS_TC_H)
#undef GOPS_TC_H
  tc::publish_sync();
  if (tid == 0) {
    umma::fence_after_sync();
    tc::issue3<8>(cx.tmem, tc::op_k_act(cx.P, TC_PLANE), tc::op_k_w(cx.W + L.o_w2, 64 * 64),
                  umma::idesc_tf32(128, 64, false, false));
    umma::commit(cx.bar);
  }
  tc::wait_mma(cx);
  float zp[MAXA];
#define GOPS_TC_L(A) tc::epi_last<A>(tl, cx.W + L.o_b2 + 16 * c, cx.W + L.o_w3, L.out, c, zp)
  GOPS_ACT_SWITCH(L.hact, GOPS_TC_L)
#undef GOPS_TC_L
#pragma unroll
  for (int a = 0; a < MAXA; ++a)
    if (a < L.out) cx.Zp[(c * MAXA + a) * 128 + r] = zp[a];
  umma::fence_before_sync();          // this sub-tile's TMEM reads before the next MMA overwrites the accumulator
  __syncthreads();
  if (tid < 128) {
#pragma unroll
    for (int a = 0; a < MAXA; ++a)
      if (a < L.out)
        Zout[a * XS + tid] = cx.W[L.o_b3 + a] + ((cx.Zp[a * 128 + tid] + cx.Zp[(MAXA + a) * 128 + tid]) +
                                                 (cx.Zp[(2 * MAXA + a) * 128 + tid] + cx.Zp[(3 * MAXA + a) * 128 + tid]));
  }
  __syncthreads();
}

}  // namespace gops
