// throughput (warp instructions per clock per SM) of the conversions / float64 operations in the epilogue of
// k_emit_tc: 16 warps per SM, 8 independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
#define N 1024
#define TP(name, decl, body, sink) __global__ void name(double* out, long long* cyc) { decl; \
  __syncthreads(); long long t0 = clock64(); _Pragma("unroll 4") for (int i = 0; i < N; ++i) { body; } long long t1 = clock64(); \
  out[blockIdx.x * blockDim.x + threadIdx.x] = sink; if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0; }
#define D8(x) x(0) x(1) x(2) x(3) x(4) x(5) x(6) x(7)
#define DECLF(j) float f##j = threadIdx.x * 1e-3f + j; double d##j = j;
#define F2F(j) d##j += (double)f##j; f##j += 1.0f;
#define F2FONLY(j) d##j = (double)f##j; f##j = __double2float_rn(d##j) + 1.0f;
#define DFMA_(j) d##j = fma(d##j, 1.0000001, 1e-9);
#define DADD_(j) d##j = d##j + 1e-9;
#define FFMA_(j) f##j = fmaf(f##j, 1.0000001f, 1e-9f);
#define SUM (d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7 + f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7)
TP(k_f2f_dadd_fadd, D8(DECLF), D8(F2F), SUM)
TP(k_dfma, D8(DECLF), D8(DFMA_), SUM)
TP(k_dadd, D8(DECLF), D8(DADD_), SUM)
TP(k_ffma, D8(DECLF), D8(FFMA_), SUM)
TP(k_f2f_roundtrip, D8(DECLF), D8(F2FONLY), SUM)
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&cyc, 64);
  long long c;
#define RUN(K, label, nops) K<<<148, 512>>>(out, cyc); K<<<148, 512>>>(out, cyc); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); \
  printf("%-28s %.3f warp-instr/clk/SM (counted ops only: %d per iteration)\n", label, 16.0 * N * 8 * nops / (double)c, nops);
  RUN(k_dfma, "DFMA", 1); RUN(k_dadd, "DADD", 1); RUN(k_ffma, "FFMA", 1);
  RUN(k_f2f_dadd_fadd, "F2F.F64.F32 + DADD + FADD", 3); RUN(k_f2f_roundtrip, "F2F.F64.F32 + F2F.F32.F64 + FADD", 3);
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}
