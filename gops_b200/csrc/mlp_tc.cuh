// tcgen05 / TMEM path of the 64-wide MLP (gops/apprfunc/mlp.py:73-77,103-111,327-329): batched inference with the
// two hidden-layer GEMMs on the 5th-generation tensor cores.
//
//   per 128-sample tile, one warpgroup (thread r <-> sample r <-> TMEM lane r):
//     X planes (hi | lo, chunk-major, see umma.cuh) <- observation rows (+ time column)
//     TMEM acc[128 x 64] = Xh.W1h^T + Xl.W1h^T + Xh.W1l^T            3xTF32, tcgen05.mma kind::tf32, M=128 N=64 K=8
//     tcgen05.ld -> + b1 -> activation -> H1 planes (hi | lo)
//     TMEM acc = H1h.W2h^T + H1l.W2h^T + H1h.W2l^T
//     tcgen05.ld -> + b2 -> activation -> output layer (thread-local dot with W3) -> squash -> global
//   weights: packed once per call into chunk-major hi / lo planes (pack_params_tc_kernel), staged by TMA bulk copy.
// Two warpgroups per CTA run independent tiles so that one's MMA / TMEM round trip overlaps the other's epilogue.
#pragma once
#include "rollout.cuh"
#include "umma.cuh"

namespace gops {

struct TcNet {
  int in, obs, out, hact, time_input, k1;          // k1 = in rounded up to 8
  int g_w1, g_b1, g_w2, g_b2, g_w3, g_b3;          // torch flat offsets
  int o_w1h, o_w1l, o_w2h, o_w2l, o_w3, o_b1, o_b2, o_b3, blob;   // packed blob offsets (floats)
  int squash;
  float half[MAXA], mid[MAXA];
};

constexpr int TC_TILE = 128;      // samples per tile = UMMA M

// WGS = warpgroups per CTA (2 when the planes fit, 1 for wide observations)
inline __host__ size_t tc_infer_smem_bytes(const TcNet& T, int WGS) {
  // 128 B header (mbarriers, TMEM slot) | weight blob | per warpgroup: X planes (2 * k1 * 128) + H1 planes (2 * 64 * 128)
  return 128 + sizeof(float) * ((size_t)T.blob + (size_t)WGS * (2 * T.k1 * TC_TILE + 2 * 64 * TC_TILE));
}

// torch-layout flat parameters -> chunk-major hi / lo planes  plane[kc][n][4] (n = output feature, kc = k / 4)
__global__ void pack_params_tc_kernel(const float* __restrict__ flat, TcNet T, float* __restrict__ blob) {
  const int n = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = t0; i < 64 * T.k1; i += n) {
    const int kc = i / 256, o = (i >> 2) & 63, k = 4 * kc + (i & 3);
    const float w = k < T.in ? flat[T.g_w1 + o * T.in + k] : 0.f;
    float hi, lo;
    umma::split(w, hi, lo);
    blob[T.o_w1h + i] = hi;
    blob[T.o_w1l + i] = lo;
  }
  for (int i = t0; i < 64 * 64; i += n) {
    const int kc = i / 256, o = (i >> 2) & 63, k = 4 * kc + (i & 3);
    float hi, lo;
    umma::split(flat[T.g_w2 + o * 64 + k], hi, lo);
    blob[T.o_w2h + i] = hi;
    blob[T.o_w2l + i] = lo;
  }
  for (int i = t0; i < T.out * 64; i += n) blob[T.o_w3 + i] = flat[T.g_w3 + i];
  for (int i = t0; i < 64; i += n) {
    blob[T.o_b1 + i] = flat[T.g_b1 + i];
    blob[T.o_b2 + i] = flat[T.g_b2 + i];
  }
  for (int i = t0; i < 4; i += n) blob[T.o_b3 + i] = i < T.out ? flat[T.g_b3 + i] : 0.f;
}

// One elected thread: D = A.B^T in 3xTF32 over `ksteps` K-steps of 8.  Planes are chunk-major with RA / 64 rows.
template <int RA>
__device__ __forceinline__ void issue_3xtf32(uint32_t d_tmem, const float* Ah, const float* Al, const float* Bh,
                                             const float* Bl, int ksteps, uint32_t idesc) {
  constexpr uint32_t LBO_A = RA * 16, LBO_B = 64 * 16, SBO = 128;
  const uint32_t ah = smem_u32(Ah), al = smem_u32(Al), bh = smem_u32(Bh), bl = smem_u32(Bl);
  uint32_t acc = 0;
  // small terms first (lo.hi, hi.lo), the dominant hi.hi last
  for (int ks = 0; ks < ksteps; ++ks, acc = 1)
    umma::mma_tf32_ss(d_tmem, umma::smem_desc(al + ks * 2 * LBO_A, LBO_A, SBO),
                      umma::smem_desc(bh + ks * 2 * LBO_B, LBO_B, SBO), idesc, acc);
  for (int ks = 0; ks < ksteps; ++ks)
    umma::mma_tf32_ss(d_tmem, umma::smem_desc(ah + ks * 2 * LBO_A, LBO_A, SBO),
                      umma::smem_desc(bl + ks * 2 * LBO_B, LBO_B, SBO), idesc, 1);
  for (int ks = 0; ks < ksteps; ++ks)
    umma::mma_tf32_ss(d_tmem, umma::smem_desc(ah + ks * 2 * LBO_A, LBO_A, SBO),
                      umma::smem_desc(bh + ks * 2 * LBO_B, LBO_B, SBO), idesc, 1);
}

// Hidden-layer epilogue of one thread (= one sample = one TMEM lane): 64 accumulator columns -> + bias -> activation
// -> hi / lo planes of the next GEMM's A operand.  16 columns per tcgen05.ld; ACT is a compile-time constant so the
// 16 independent elements interleave.
template <int ACT>
__device__ __forceinline__ void tc_epilogue_hidden(uint32_t ld_tmem, const float* __restrict__ bias,
                                                   float* __restrict__ Hh, float* __restrict__ Hl, int r) {
#pragma unroll 1
  for (int cg = 0; cg < 4; ++cg) {
    float v[16];
    umma::tmem_ld16(ld_tmem + cg * 16, v);
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
      const float4 b = *reinterpret_cast<const float4*>(bias + cg * 16 + c4 * 4);
      float4 h, l;
      umma::split(act_fwd_t<ACT>(v[4 * c4 + 0] + b.x), h.x, l.x);
      umma::split(act_fwd_t<ACT>(v[4 * c4 + 1] + b.y), h.y, l.y);
      umma::split(act_fwd_t<ACT>(v[4 * c4 + 2] + b.z), h.z, l.z);
      umma::split(act_fwd_t<ACT>(v[4 * c4 + 3] + b.w), h.w, l.w);
      reinterpret_cast<float4*>(Hh)[(cg * 4 + c4) * TC_TILE + r] = h;
      reinterpret_cast<float4*>(Hl)[(cg * 4 + c4) * TC_TILE + r] = l;
    }
  }
}
// Last hidden layer + output layer: z[a] += sum_j W3[a][j] * act(acc[j] + b2[j])   (thread-local)
template <int ACT>
__device__ __forceinline__ void tc_epilogue_out(uint32_t ld_tmem, const float* __restrict__ bias,
                                                const float* __restrict__ W3, int out, float* z) {
#pragma unroll 1
  for (int cg = 0; cg < 4; ++cg) {
    float v[16];
    umma::tmem_ld16(ld_tmem + cg * 16, v);
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
      const float4 b = *reinterpret_cast<const float4*>(bias + cg * 16 + c4 * 4);
      const float h0 = act_fwd_t<ACT>(v[4 * c4 + 0] + b.x), h1 = act_fwd_t<ACT>(v[4 * c4 + 1] + b.y);
      const float h2 = act_fwd_t<ACT>(v[4 * c4 + 2] + b.z), h3 = act_fwd_t<ACT>(v[4 * c4 + 3] + b.w);
#pragma unroll
      for (int a = 0; a < MAXA; ++a)
        if (a < out) {
          const float4 w = *reinterpret_cast<const float4*>(W3 + a * 64 + cg * 16 + c4 * 4);
          z[a] = fmaf(w.w, h3, fmaf(w.z, h2, fmaf(w.y, h1, fmaf(w.x, h0, z[a]))));
        }
    }
  }
}

template <int TC_WGS>
__global__ void __launch_bounds__(128 * TC_WGS, 1)
    mlp_infer_tc_kernel(const __grid_constant__ TcNet T, const float* __restrict__ blob, const float* __restrict__ obs,
                        long long B, float virtual_t, float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* wbar = reinterpret_cast<uint64_t*>(smem_raw);            // weights landed
  uint64_t* gbar = reinterpret_cast<uint64_t*>(smem_raw + 16);       // [TC_WGS]: MMA group of the warpgroup done
  uint32_t* tslot = reinterpret_cast<uint32_t*>(smem_raw + 64);      // TMEM base address
  float* W = reinterpret_cast<float*>(smem_raw + 128);
  const int tid = threadIdx.x, wg = tid >> 7, r = tid & 127, warp = tid >> 5;
  float* Xh = W + T.blob + (size_t)wg * (2 * T.k1 * TC_TILE + 2 * 64 * TC_TILE);
  float* Xl = Xh + T.k1 * TC_TILE;
  float* Hh = Xl + T.k1 * TC_TILE;
  float* Hl = Hh + 64 * TC_TILE;
  constexpr uint32_t NCOLS = 64 * TC_WGS;                            // power of two >= 32

  if (tid == 0) {
    mbar_init(wbar, 1);
    for (int g = 0; g < TC_WGS; ++g) mbar_init(gbar + g, 1);
    fence_mbar_init();
  }
  if (warp == 0) umma::tmem_alloc(tslot, NCOLS);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem_base = *tslot;
  if (tid == 0) {
    const uint32_t bytes = (uint32_t)T.blob * 4u;
    mbar_expect_tx(wbar, bytes);
    for (uint32_t off = 0; off < bytes; off += 32768u) {
      const uint32_t nb = bytes - off < 32768u ? bytes - off : 32768u;
      tma_bulk_g2s(reinterpret_cast<char*>(W) + off, reinterpret_cast<const char*>(blob) + off, nb, wbar);
    }
  }
  mbar_wait(wbar, 0);

  const uint32_t d_tmem = tmem_base + wg * 64;                               // accumulator columns of this warpgroup
  const uint32_t ld_tmem = d_tmem + ((uint32_t)(32 * (warp & 3)) << 16);     // + this warp's lane quarter
  constexpr uint32_t IDESC = umma::idesc_tf32(128, 64, false, false);
  const long long n_tiles = (B + TC_TILE - 1) / TC_TILE;
  uint32_t phase = 0;
  for (long long tile = (long long)blockIdx.x * TC_WGS + wg; tile < n_tiles; tile += (long long)gridDim.x * TC_WGS) {
    const long long s = tile * TC_TILE + r;
    const bool valid = s < B;
    // ---- observation row -> X planes
    for (int kc = 0; kc < T.k1 / 4; ++kc) {
      float4 h, l;
      float x[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int f = 4 * kc + q;
        x[q] = !valid ? 0.f : f < T.obs ? obs[s * T.obs + f] : (f < T.in ? virtual_t : 0.f);
      }
      umma::split(x[0], h.x, l.x); umma::split(x[1], h.y, l.y);
      umma::split(x[2], h.z, l.z); umma::split(x[3], h.w, l.w);
      reinterpret_cast<float4*>(Xh)[kc * TC_TILE + r] = h;
      reinterpret_cast<float4*>(Xl)[kc * TC_TILE + r] = l;
    }
    fence_proxy_async();            // generic-proxy smem writes -> visible to the tensor core (async proxy)
    umma::fence_before_sync();
    umma::wg_sync(wg);
    if (r == 0) {
      umma::fence_after_sync();
      issue_3xtf32<TC_TILE>(d_tmem, Xh, Xl, W + T.o_w1h, W + T.o_w1l, T.k1 / 8, IDESC);
      umma::commit(gbar + wg);
    }
    mbar_wait(gbar + wg, phase);
    phase ^= 1;
    umma::fence_after_sync();
    // ---- layer-1 epilogue: + b1, activation, split, H1 planes
#define GOPS_TC_EPI1(A) tc_epilogue_hidden<A>(ld_tmem, W + T.o_b1, Hh, Hl, r)
    GOPS_ACT_SWITCH(T.hact, GOPS_TC_EPI1)
#undef GOPS_TC_EPI1
    fence_proxy_async();
    umma::fence_before_sync();      // orders the tcgen05.ld above before the MMA that overwrites the accumulator
    umma::wg_sync(wg);
    if (r == 0) {
      umma::fence_after_sync();
      issue_3xtf32<TC_TILE>(d_tmem, Hh, Hl, W + T.o_w2h, W + T.o_w2l, 8, IDESC);
      umma::commit(gbar + wg);
    }
    mbar_wait(gbar + wg, phase);
    phase ^= 1;
    umma::fence_after_sync();
    // ---- layer-2 epilogue + output layer (thread-local)
    float z[MAXA];
#pragma unroll
    for (int a = 0; a < MAXA; ++a) z[a] = a < T.out ? W[T.o_b3 + a] : 0.f;
#define GOPS_TC_EPI2(A) tc_epilogue_out<A>(ld_tmem, W + T.o_b2, W + T.o_w3, T.out, z)
    GOPS_ACT_SWITCH(T.hact, GOPS_TC_EPI2)
#undef GOPS_TC_EPI2
    if (valid) {
#pragma unroll
      for (int a = 0; a < MAXA; ++a)
        if (a < T.out) {
          float y = z[a];
          if (T.squash) y = __fadd_rn(__fmul_rn(T.half[a], tanhf(y)), T.mid[a]);
          out[s * T.out + a] = y;
        }
    }
    umma::fence_before_sync();      // this tile's TMEM reads before the next tile's first MMA
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem_base, NCOLS);
}

}  // namespace gops
