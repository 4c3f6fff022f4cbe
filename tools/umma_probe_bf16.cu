// Layout probe for tcgen05.mma kind::f16 with BF16 operands (companion of umma_probe.cu; same method):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I gops_b200/csrc -I include tools/umma_probe_bf16.cu -o build/umma_probe_bf16
// The probed buffer holds, as bf16, one byte of every 16-bit word's own index per pass (3 passes); multiplying by a one-hot
// operand (K = 16) copies the probed operand into the accumulator as the tensor core addresses it.
#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "umma.cuh"

using namespace gops;

struct Probe {
  int kind;            // 0: probe A (B one-hot K-major), 1: probe B (A one-hot K-major)
  int a_mn, b_mn;
  uint32_t layout, lbo, sbo;
  int M, N;
};

__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(d),
      "l"(a), "l"(b), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0)
      : "memory");
}
__device__ __forceinline__ uint64_t desc_l(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return umma::smem_desc(saddr, lbo, sbo) | ((uint64_t)layout << 61);
}

__global__ void __launch_bounds__(128, 1) probe_kernel(Probe p, int words, int pass, float* out /* [128][64] */) {
  extern __shared__ __align__(1024) unsigned char raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(raw);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(raw + 16);
  __nv_bfloat16* probed = reinterpret_cast<__nv_bfloat16*>(raw + (1024 - (smem_u32(raw) & 1023u)) % 1024 + 1024);
  __nv_bfloat16* onehot = probed + words;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < words; i += blockDim.x) probed[i] = __float2bfloat16((float)((i >> (8 * pass)) & 255));
  const int R = p.kind == 0 ? 64 : 128;      // one-hot K-major no-swizzle: [2 chunks][R rows][8], (r, k) = (r == k)
  for (int i = tid; i < 2 * R * 8; i += blockDim.x) {
    const int ch = i / (R * 8), r = (i / 8) % R, e = i & 7;
    onehot[i] = __float2bfloat16(r == 8 * ch + e ? 1.f : 0.f);
  }
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) umma::tmem_alloc(tslot, 64);
  fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tslot;
  if (tid == 0) {
    const uint32_t idesc = idesc_bf16(p.M, p.N, p.a_mn != 0, p.b_mn != 0);
    if (p.kind == 0)
      mma_bf16_ss(tmem, desc_l(smem_u32(probed), p.lbo, p.sbo, p.layout), umma::smem_desc(smem_u32(onehot), 64 * 16, 128),
                  idesc, 0);
    else
      mma_bf16_ss(tmem, umma::smem_desc(smem_u32(onehot), 128 * 16, 128), desc_l(smem_u32(probed), p.lbo, p.sbo, p.layout),
                  idesc, 0);
    umma::commit(bar);
  }
  mbar_wait(bar, 0);
  umma::fence_after_sync();
  const uint32_t tl = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
  for (int cg = 0; cg < 4; ++cg) {
    float v[16];
    umma::tmem_ld16(tl + cg * 16, v);
    for (int e = 0; e < 16; ++e) out[tid * 64 + cg * 16 + e] = v[e];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 64);
}

static void show(const char* name, const Probe& p, int words) {
  printf("== %s  (a_mn %d b_mn %d layout %u LBO %u SBO %u M %d N %d)\n", name, p.a_mn, p.b_mn, p.layout, p.lbo, p.sbo, p.M, p.N);
  float* dout;
  cudaMalloc(&dout, 128 * 64 * sizeof(float));
  std::vector<float> h(128 * 64);
  std::vector<int> idx(128 * 64, 0);
  const size_t smem = 2048 + (size_t)words * 2 + 2 * 128 * 8 * 2 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int pass = 0; pass < 2; ++pass) {
    cudaMemset(dout, 0, 128 * 64 * sizeof(float));
    probe_kernel<<<1, 128, smem>>>(p, words, pass, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(h.data(), dout, 128 * 64 * sizeof(float), cudaMemcpyDeviceToHost);
    for (int i = 0; i < 128 * 64; ++i) idx[i] += ((int)h[i]) << (8 * pass);
  }
  cudaFree(dout);
  if (p.kind == 0) {
    const int ms[] = {0, 1, 2, 7, 8, 9, 16, 31, 32, 63, 64, 65, 127};
    for (int m : ms) {
      printf("  m %3d:", m);
      for (int k = 0; k < 16; ++k) printf(" %5d", idx[m * 64 + k]);
      printf("\n");
    }
  } else {
    const int ns[] = {0, 1, 2, 7, 8, 9, 16, 31, 32, 63};
    for (int n : ns) {
      if (n >= p.N) continue;
      printf("  n %3d:", n);
      for (int k = 0; k < 16; ++k) printf(" %5d", idx[k * 64 + n]);
      printf("\n");
    }
  }
}

int main() {
  const int W = 32768;   // 16-bit words (64 KB)
  // plane[k/8][row][8]: K-major no-swizzle, 128 rows: LBO = 128 * 16, SBO = 128
  show("A K-major no-swizzle 128 rows", Probe{0, 0, 0, 0, 2048, 128, 128, 64}, W);
  // the same buffer read transposed: mn = k-index of the plane, k = row
  show("A MN-major no-swizzle LBO 128 SBO 2048", Probe{0, 1, 0, 0, 128, 2048, 128, 64}, W);
  show("A MN-major no-swizzle LBO 2048 SBO 128", Probe{0, 1, 0, 0, 2048, 128, 128, 64}, W);
  show("B K-major no-swizzle 64 rows", Probe{1, 0, 0, 0, 1024, 128, 128, 64}, W);
  show("B MN-major no-swizzle LBO 128 SBO 1024 N 64", Probe{1, 0, 1, 0, 128, 1024, 128, 64}, W);
  show("B MN-major no-swizzle LBO 128 SBO 2048 N 16", Probe{1, 0, 1, 0, 128, 2048, 128, 16}, W);
  return 0;
}
