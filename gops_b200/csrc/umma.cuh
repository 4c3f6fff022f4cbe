// tcgen05 (5th-generation tensor core) primitives for sm_100a: shared-memory matrix descriptors, the TF32
// instruction descriptor, single-thread MMA issue, commit -> mbarrier, TMEM allocation and TMEM -> register loads.
//
// Operand layout used throughout ("chunk-major", the no-swizzle canonical UMMA layout with LBO = R*16, SBO = 128):
//   plane[kc][r][4]   r = 0..R-1 rows (samples for activations, output features for weights), kc = k / 4
// i.e. core matrix (8 rows x 16 bytes) (r/8, kc) starts at byte  kc * (R*16) + (r/8) * 128  and row r%8 of it at +16*(r%8).
//   K-major view  (M/N = r, K = k):  ((8,m),(4,2)) : ((16 B, SBO = 128), (4 B, LBO = R*16))
//   MN-major view (M/N = k, K = r):  ((4,1,m),(8,k)) : ((4 B, -, SBO = R*16), (16 B, LBO = 128))   -- the transposed matrix
// so one buffer serves X.W^T (activations K-major) and, transposed, the weight-gradient contraction over samples.
// A thread that owns row r writes float4 chunks at  kc * (R*16) + r*16: a warp covers 512 contiguous bytes (no conflicts).
#pragma once
#include "common.cuh"

namespace gops {
namespace umma {

// 64-bit shared-memory matrix descriptor (no swizzle): start address, leading / stride byte offsets (>> 4), version 1
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

// 32-bit instruction descriptor, kind::tf32, FP32 accumulate: c_format F32 (bits 4-5 = 1), a/b format TF32 (= 2,
// bits 7-9 / 10-12), a/b major (bit 15 / 16: 0 = K-major, 1 = MN-major), N >> 3 (bits 17-22), M >> 4 (bits 24-28)
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] . B[smem]^T, issued by ONE thread; K = 8 TF32 elements per instruction
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
      : "memory");
}
// mbarrier arrive when every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// one lane of a converged warp (warp-uniform MMA issue: `if (warp_uniform_cond) if (elect_one()) { ... }`)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
  return pred != 0;
}

// TMEM allocation (one full warp); the base address (lane << 16 | column) is written to *slot
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// this thread's TMEM lane (= 32 * (warp % 4) + lane), 16 consecutive 32-bit columns starting at taddr's column
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// registers -> this thread's TMEM lane, 16 consecutive columns (TMEM as per-sample scratch: lane = sample)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16};\n" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// x = hi + lo exactly, hi on the TF32 grid (round to nearest, ties away); the tensor core truncates lo
__device__ __forceinline__ void split(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
  lo = x - hi;
}

// 128-thread named barrier of one warpgroup (ids 1..; id 0 is __syncthreads)
__device__ __forceinline__ void wg_sync(int wg) { asm volatile("bar.sync %0, 128;" ::"r"(1 + wg) : "memory"); }

}  // namespace umma
}  // namespace gops
