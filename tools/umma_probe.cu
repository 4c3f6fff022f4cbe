// Layout probe for tcgen05.mma kind::tf32 operands on sm_100a (standalone; build + run on the GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I gops_b200/csrc -I include tools/umma_probe.cu -o build/umma_probe
// Shared memory is filled with the float value of each word's own index (0..N-1, exact in TF32 below 2048; larger
// buffers use index mod 2048 plus a second pass with index / 2048).  Multiplying by a one-hot operand makes the
// accumulator a copy of the other operand AS THE TENSOR CORE ADDRESSES IT, i.e. D[m][n] = word index that the
// descriptor maps to logical element (m, k = n) (A probes) or (n = column, k = m) (B probes).  Printed tables give the
// address law of a descriptor (layout type, majorness, LBO, SBO) without guessing from documentation.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "umma.cuh"

using namespace gops;

struct Probe {
  int kind;            // 0: probe A (B one-hot K-major), 1: probe B (A one-hot K-major), 2: A from TMEM, 3: M=64 lanes
  int a_mn, b_mn;      // majorness bits of the instruction descriptor
  uint32_t layout;     // descriptor layout type of the PROBED operand (bits 61..63)
  uint32_t lbo, sbo;   // bytes, probed operand
  int M, N;
  uint32_t start_off;  // byte offset of the probed operand's start address inside its buffer
};

__device__ __forceinline__ uint64_t desc_l(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  return umma::smem_desc(saddr, lbo, sbo) | ((uint64_t)layout << 61);
}

// one-hot K-major no-swizzle operand with R rows (chunk-major: [2 chunks][R rows][4]): element (r, k) = (r == k)
__device__ void fill_onehot(float* buf, int R) {
  for (int i = threadIdx.x; i < 2 * R * 4; i += blockDim.x) {
    const int ch = i / (R * 4), r = (i / 4) % R, e = i & 3, k = 4 * ch + e;
    buf[i] = (r == k) ? 1.f : 0.f;
  }
}

__global__ void __launch_bounds__(128, 1) probe_kernel(Probe p, int words, int pass, float* out /* [128][64] */) {
  extern __shared__ __align__(1024) unsigned char raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(raw);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(raw + 16);
  // 1024-byte aligned in the shared window (swizzle atoms repeat every 1024 bytes)
  float* probed = reinterpret_cast<float*>(raw + (1024 - (smem_u32(raw) & 1023u)) % 1024 + 1024);
  float* onehot = probed + words;                                        // after the probed buffer
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < words; i += blockDim.x) probed[i] = pass == 0 ? (float)(i & 2047) : (float)(i >> 11);
  fill_onehot(onehot, p.kind == 0 || p.kind == 2 ? 64 : 128);
  if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  if (warp == 0) umma::tmem_alloc(tslot, 256);
  fence_proxy_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tslot;
  const uint32_t tl = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
  if (p.kind == 2) {   // A operand in TMEM columns 128..135: lane m, column k holds m * 8 + k
    float v[16];
    for (int k = 0; k < 16; ++k) v[k] = (float)((tid * 8 + k) & 2047);
    umma::tmem_st16(tl + 128, v);
    umma::tmem_wait_st();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
  }
  if (tid == 0) {
    const uint32_t idesc = umma::idesc_tf32(p.M, p.N, p.a_mn != 0, p.b_mn != 0);
    if (p.kind == 0) {
      const uint64_t a = desc_l(smem_u32(probed) + p.start_off, p.lbo, p.sbo, p.layout);
      const uint64_t b = umma::smem_desc(smem_u32(onehot), 64 * 16, 128);          // K-major one-hot, 64 rows
      umma::mma_tf32_ss(tmem, a, b, idesc, 0);
    } else if (p.kind == 1) {
      const uint64_t a = umma::smem_desc(smem_u32(onehot), 128 * 16, 128);         // K-major one-hot, 128 rows
      const uint64_t b = desc_l(smem_u32(probed) + p.start_off, p.lbo, p.sbo, p.layout);
      umma::mma_tf32_ss(tmem, a, b, idesc, 0);
    } else if (p.kind == 2) {
      const uint64_t b = umma::smem_desc(smem_u32(onehot), 64 * 16, 128);
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(tmem),
          "r"(tmem + 128), "l"(b), "r"(idesc), "r"(0), "r"(0), "r"(0), "r"(0), "r"(0)
          : "memory");
    } else {             // M = 64: A = probed K-major no-swizzle (64 rows), B one-hot
      const uint64_t a = umma::smem_desc(smem_u32(probed), 64 * 16, 128);
      const uint64_t b = umma::smem_desc(smem_u32(onehot), 64 * 16, 128);
      umma::mma_tf32_ss(tmem, a, b, idesc, 0);
    }
    umma::commit(bar);
  }
  mbar_wait(bar, 0);
  umma::fence_after_sync();
  for (int cg = 0; cg < 4; ++cg) {
    float v[16];
    umma::tmem_ld16(tl + cg * 16, v);
    for (int e = 0; e < 16; ++e) out[tid * 64 + cg * 16 + e] = v[e];
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

static std::vector<int> run(const Probe& p, int words) {
  float* dout;
  cudaMalloc(&dout, 128 * 64 * sizeof(float));
  std::vector<float> h0(128 * 64), h1(128 * 64);
  const size_t smem = 2048 + (size_t)words * 4 + 2 * 128 * 4 * 4 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int pass = 0; pass < 2; ++pass) {
    cudaMemset(dout, 0xff, 128 * 64 * sizeof(float));
    probe_kernel<<<1, 128, smem>>>(p, words, pass, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("  CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
    cudaMemcpy(pass == 0 ? h0.data() : h1.data(), dout, 128 * 64 * sizeof(float), cudaMemcpyDeviceToHost);
  }
  cudaFree(dout);
  std::vector<int> idx(128 * 64);
  for (int i = 0; i < 128 * 64; ++i) idx[i] = (int)h0[i] + 2048 * (int)h1[i];
  return idx;
}

static void show(const char* name, const Probe& p, int words, bool probe_b) {
  printf("== %s  (kind %d a_mn %d b_mn %d layout %u LBO %u SBO %u M %d N %d start+%u)\n", name, p.kind, p.a_mn, p.b_mn,
         p.layout, p.lbo, p.sbo, p.M, p.N, p.start_off);
  std::vector<int> idx = run(p, words);
  if (!probe_b) {   // D[m][k]: word index of A element (m, k)
    const int ms[] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 16, 31, 32, 33, 63, 64, 65, 96, 127};
    for (int m : ms) {
      printf("  m %3d:", m);
      for (int k = 0; k < 8; ++k) printf(" %6d", idx[m * 64 + k]);
      printf("\n");
    }
  } else {          // D[m = k][n]: word index of B element (n, k)
    const int ns[] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 16, 31, 32, 33, 63};
    for (int n : ns) {
      if (n >= p.N) continue;
      printf("  n %3d:", n);
      for (int k = 0; k < 8; ++k) printf(" %6d", idx[k * 64 + n]);
      printf("\n");
    }
  }
}

int main() {
  const int W = 16384;   // 64 KB probed buffer
  // sanity: the layouts the product kernels use (K-major, no swizzle)
  show("A K-major no-swizzle, 128 rows (plane[k/4][row][4])", Probe{0, 0, 0, 0, 2048, 128, 128, 64, 0}, W, false);
  show("B K-major no-swizzle, 64 rows", Probe{1, 0, 0, 0, 1024, 128, 128, 64, 0}, W, true);
  // transposed reads, no swizzle (expected: not supported for 32-bit operands)
  show("A MN-major no-swizzle LBO 128 SBO 2048", Probe{0, 1, 0, 0, 128, 2048, 128, 64, 0}, W, false);
  // transposed reads, SWIZZLE_128B_BASE32B (layout type 1): vary LBO / SBO to see which one strides what
  show("A MN-major 128B_BASE32B LBO 4096 SBO 512", Probe{0, 1, 0, 1, 4096, 512, 128, 64, 0}, W, false);
  show("A MN-major 128B_BASE32B LBO 8192 SBO 1024", Probe{0, 1, 0, 1, 8192, 1024, 128, 64, 0}, W, false);
  show("A MN-major 128B_BASE32B LBO 4096 SBO 512 start+512", Probe{0, 1, 0, 1, 4096, 512, 128, 64, 512}, W, false);
  show("B MN-major 128B_BASE32B LBO 4096 SBO 512 N 64", Probe{1, 0, 1, 1, 4096, 512, 128, 64, 0}, W, true);
  show("B MN-major 128B_BASE32B LBO 4096 SBO 512 N 16", Probe{1, 0, 1, 1, 4096, 512, 128, 16, 0}, W, true);
  // other swizzles, transposed
  show("A MN-major SWIZZLE_128B (type 2) LBO 4096 SBO 1024", Probe{0, 1, 0, 2, 4096, 1024, 128, 64, 0}, W, false);
  // K-major with 128B swizzle (for reference)
  show("A K-major SWIZZLE_128B (type 2) SBO 1024", Probe{0, 0, 0, 2, 16, 1024, 128, 64, 0}, W, false);
  // A operand from TMEM
  show("A from TMEM (lane m, column k = m*8+k)", Probe{2, 0, 0, 0, 0, 0, 128, 64, 0}, W, false);
  // M = 64: where do the 64 rows land in the 128 TMEM lanes?
  show("M = 64 (rows -> lanes)", Probe{3, 0, 0, 0, 0, 0, 64, 64, 0}, W, false);
  return 0;
}
