// Probe for the tcgen05 building blocks used by the dense K >= 128 recursion kernel: shared-memory
// (K-major, SWIZZLE_128B) descriptors, instruction descriptor, TMEM alloc / ld, commit -> mbarrier.
// D[128 x 256] = A[128 x 256] . B[256 x 256]^T, bf16 inputs, fp32 accumulate, checked on the host.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#define M 128
#define N 256
#define KD 256

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;                 // LBO (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;       // SBO: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                 // version 1 (Blackwell)
  d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
  return d;
}
// byte offset of element (row r, column k) of a K-major operand with `rows` rows, 64-column slabs
__device__ __forceinline__ uint32_t sw_off(int r, int k, int rows) {
  const int kb = k >> 6, kk = k & 63;
  return (uint32_t)kb * rows * 128 + (r >> 3) * 1024 + (r & 7) * 128 + ((((kk >> 3) ^ (r & 7)) << 4)) + (kk & 7) * 2;
}

__global__ void __launch_bounds__(128) k_probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint8_t* sA = sm;                          // 128 x 256 bf16 = 64 KB
  uint8_t* sB = sm + M * KD * 2;             // 256 x 256 bf16 = 128 KB
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, wp = tid >> 5;
  for (int i = tid; i < M * KD; i += 128) { const int r = i / KD, k = i % KD; *(__nv_bfloat16*)(sA + sw_off(r, k, M)) = A[i]; }
  for (int i = tid; i < N * KD; i += 128) { const int r = i / KD, k = i % KD; *(__nv_bfloat16*)(sB + sw_off(r, k, N)) = B[i]; }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (wp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int ks = 0; ks < KD / 16; ++ks) {
      const uint64_t da = make_desc(smem_u32(sA) + (ks >> 2) * (M * 128) + (ks & 3) * 32);
      const uint64_t db = make_desc(smem_u32(sB) + (ks >> 2) * (N * 128) + (ks & 3) * 32);
      const uint32_t acc = ks > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // thread = row (TMEM lane); 8 chunks of 32 columns
  for (int c = 0; c < N / 32; ++c) {
    uint32_t v[32];
    const uint32_t ta = tm + ((uint32_t)(wp * 32) << 16) + c * 32;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                   "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                   "=r"(v[30]), "=r"(v[31]) : "r"(ta) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) D[(size_t)tid * N + c * 32 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (wp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256) : "memory");
}

int main() {
  std::vector<__nv_bfloat16> hA(M * KD), hB(N * KD);
  std::vector<float> fA(M * KD), fB(N * KD);
  srand(1);
  for (int i = 0; i < M * KD; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hA[i] = __float2bfloat16(v); fA[i] = __bfloat162float(hA[i]); }
  for (int i = 0; i < N * KD; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, M * KD * 2); cudaMalloc(&dB, N * KD * 2); cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, hA.data(), M * KD * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), N * KD * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, M * N * 4);
  const int smem = M * KD * 2 + N * KD * 2 + 1024;
  cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_probe<<<1, 128, smem>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  std::vector<float> hD(M * N);
  cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < KD; ++k) s += (double)fA[m * KD + k] * fB[n * KD + k];
      const double er = fabs(s - hD[m * N + n]);
      if (er > maxerr) maxerr = er;
      if (er > 1e-2 && bad < 5) { printf("mismatch m=%d n=%d ref=%f got=%f\n", m, n, s, hD[m * N + n]); ++bad; }
    }
  printf("max abs err %.3e (%s)\n", maxerr, maxerr < 1e-2 ? "OK" : "FAIL");
  return 0;
}
