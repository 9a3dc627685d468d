// Do the FP64 pipe and the tensor pipe (tcgen05.mma) run concurrently?  One CTA per SM: thread 256 issues
// a stream of M128 N64 K16 float16 MMAs (operands: whatever is in shared memory) while warps 0-7 time a
// DFMA / FFMA loop; each is also timed alone.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../pysvihmm_b200/csrc/dense.cuh"
#define NIT 4096
template <int MODE>   // bit 0: MMA stream runs, bit 1: FP64 loop (else FP32 loop)
__global__ void __launch_bounds__(288) k(double* out, long long* cyc, int nmma) {
  extern __shared__ __align__(1024) uint8_t raw[];
  uint8_t* sm = raw + ((1024u - (dn_smem(raw) & 1023u)) & 1023u);
  __shared__ uint32_t tmem_base; __shared__ __align__(8) unsigned long long bar;
  const int tid = threadIdx.x;
  for (int i = tid; i < 48 * 1024 / 4; i += 288) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;   // 1.0 halves
  if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dn_smem(&bar)) : "memory"); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (tid < 32) { asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dn_smem(&tmem_base)), "r"(512) : "memory");
                  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;
  long long t0 = clock64(), t1 = t0;
  if (tid == 256) {
    if (MODE & 1) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t da = dn_desc(dn_smem(sm)), db = dn_desc(dn_smem(sm + 32768));
      for (int i = 0; i < nmma; ++i)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tm + (uint32_t)(i & 7) * 64), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(dn_smem(&bar)) : "memory");
      unsigned ok = 0;
      while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(dn_smem(&bar)), "r"(0) : "memory");
      t1 = clock64();
      if (blockIdx.x == 0) cyc[1] = t1 - t0;
    }
  } else if (tid < 256) {
    double d[8]; float f[8];
    for (int j = 0; j < 8; ++j) { d[j] = tid * 1e-3 + j; f[j] = tid * 1e-3f + j; }
#pragma unroll 2
    for (int i = 0; i < NIT; ++i) {
      if (MODE & 2) { _Pragma("unroll") for (int j = 0; j < 8; ++j) d[j] = fma(d[j], 1.0000001, 1e-9); }
      else { _Pragma("unroll") for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], 1.0000001f, 1e-9f); }
    }
    t1 = clock64();
    double s = 0; for (int j = 0; j < 8; ++j) s += d[j] + f[j];
    out[blockIdx.x * 256 + tid] = s;
    if (tid == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}
template <int MODE> void run(const char* label, double* out, long long* cyc, int nmma) {
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  long long c[2] = {0, 0};
  for (int r = 0; r < 2; ++r) { cudaMemset(cyc, 0, 16); k<MODE><<<148, 288, 100 * 1024>>>(out, cyc, nmma); cudaDeviceSynchronize(); }
  cudaMemcpy(c, cyc, 16, cudaMemcpyDeviceToHost);
  printf("%-44s math loop %8lld cycles (%.2f warp-instr/clk/SM)   MMA stream %8lld cycles (%.1f cycles per MMA)  %s\n", label, c[0],
         8.0 * NIT * 8 / (double)c[0], c[1], nmma ? c[1] / (double)nmma : 0.0, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 148 * 256 * 8); cudaMalloc(&cyc, 64);
  run<2>("DFMA alone", out, cyc, 0);
  run<3>("DFMA + 4000 MMAs (M128 N64 K16 f16)", out, cyc, 4000);
  run<0>("FFMA alone", out, cyc, 0);
  run<1>("FFMA + 4000 MMAs", out, cyc, 4000);
  run<3>("DFMA + 1000 MMAs", out, cyc, 1000);
}
