// float64 latencies on B200: dependent DFMA / DADD / division / log / exp chains (one warp).
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
#define BENCH(name, body) __global__ void name(double* out, long long* cyc) { double a = 1.0 + threadIdx.x * 1e-3, b = 1.000001; \
  long long t0 = clock64(); _Pragma("unroll 8") for (int i = 0; i < N; ++i) { body; } long long t1 = clock64(); \
  out[threadIdx.x] = a; if (threadIdx.x == 0) cyc[0] = t1 - t0; }
BENCH(k_dfma, a = fma(a, b, 1e-9))
BENCH(k_dadd, a = a + b)
BENCH(k_dmul, a = a * b)
BENCH(k_ddiv, a = b / a + 1.0)
BENCH(k_dlog, a = log(a) + 2.0)
BENCH(k_dexp, a = exp(a * 1e-3))
BENCH(k_drcp, a = __drcp_rn(a) + 1.0)
BENCH(k_dsqrt, a = sqrt(a) + 1.0)
__global__ void k_ldg_chase(double* out, long long* cyc, const int* nxt) { int p = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) p = __ldcg(nxt + p);
  long long t1 = clock64(); out[threadIdx.x] = p; if (threadIdx.x == 0) cyc[0] = t1 - t0; }
int main() {
  double* out; long long* cyc; int* nxt; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64); cudaMalloc(&nxt, 1 << 20);
  int* h = new int[1 << 18]; for (int i = 0; i < (1 << 18); ++i) h[i] = (i * 97 + 33) & ((1 << 18) - 1);
  cudaMemcpy(nxt, h, 1 << 20, cudaMemcpyHostToDevice);
  long long c;
#define RUN(K, label) K<<<1, 32>>>(out, cyc); K<<<1, 32>>>(out, cyc); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost); printf("%-12s %.1f cycles/op (dependent chain, incl. +1 add where noted)\n", label, (double)c / N);
  RUN(k_dfma, "DFMA"); RUN(k_dadd, "DADD"); RUN(k_dmul, "DMUL"); RUN(k_ddiv, "DDIV+DADD"); RUN(k_dlog, "log+DADD"); RUN(k_dexp, "exp(+DMUL)");
  RUN(k_drcp, "drcp+DADD"); RUN(k_dsqrt, "sqrt+DADD");
  k_ldg_chase<<<1, 32>>>(out, cyc, nxt); k_ldg_chase<<<1, 32>>>(out, cyc, nxt); cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-12s %.1f cycles/op\n", "LDG.cg chase (L2)", (double)c / N);
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
}
