// Microbenchmarks of the primitives the E-step kernels lean on (B200, sm_100a):
// SHFL throughput/latency, DFMA throughput, LDS broadcast latency, STS->LDS round trip, FFMA chain.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu && ./ubench
#include <cstdio>
#include <cuda_runtime.h>

#define N 4096
__global__ void k_shfl_tput(float* out, long long* cyc) {          // independent shuffles
  float v = threadIdx.x, a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N / 16; ++i) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      a0 += __shfl_sync(0xffffffffu, v, j, 16); a1 += __shfl_sync(0xffffffffu, v, j + 1, 16);
      a2 += __shfl_sync(0xffffffffu, v, j + 2, 16); a3 += __shfl_sync(0xffffffffu, v, j + 3, 16);
    }
    v += 1.f;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_shfl_lat(float* out, long long* cyc) {           // dependent shuffles
  float v = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) v = __shfl_sync(0xffffffffu, v, (i + 1) & 15, 16) + 1.f;
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_dfma(double* out, long long* cyc) {              // 8 independent DFMA chains
  double a[8]; for (int j = 0; j < 8; ++j) a[j] = threadIdx.x + j;
  const double m = 1.0000001, c = 1e-9;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N / 8; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = fma(a[j], m, c);
  }
  long long t1 = clock64();
  double s = 0; for (int j = 0; j < 8; ++j) s += a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_ffma(float* out, long long* cyc) {               // 8 independent FFMA chains
  float a[8]; for (int j = 0; j < 8; ++j) a[j] = threadIdx.x + j;
  const float m = 1.0000001f, c = 1e-9f;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N / 8; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = fmaf(a[j], m, c);
  }
  long long t1 = clock64();
  float s = 0; for (int j = 0; j < 8; ++j) s += a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_ffma_chain(float* out, long long* cyc) {         // one dependent FFMA chain
  float a = threadIdx.x; const float m = 1.0000001f, c = 1e-9f;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) a = fmaf(a, m, c);
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// the smem-broadcast matvec step: STS own value, 4x LDS.128 of the group's 16 values, 16 FFMA
__global__ void k_sts_lds_step(float* out, long long* cyc) {
  __shared__ __align__(16) float buf[2][8][32];   // [parity][warp][2 groups x 16]
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5, g = lane >> 4;
  float col[16]; for (int i = 0; i < 16; ++i) col[i] = 0.0625f + 1e-3f * ((lane + i) & 7);
  float v = 1.f + lane * 0.01f;
  long long t0 = clock64();
#pragma unroll 2
  for (int i = 0; i < N; ++i) {
    float* b = buf[i & 1][wp];
    b[lane] = v;
    __syncwarp();
    const float4* p = reinterpret_cast<const float4*>(b + 16 * g);
    const float4 x0 = p[0], x1 = p[1], x2 = p[2], x3 = p[3];
    float a0 = x0.x * col[0], a1 = x1.x * col[4], a2 = x2.x * col[8], a3 = x3.x * col[12];
    a0 = fmaf(x0.y, col[1], a0); a1 = fmaf(x1.y, col[5], a1); a2 = fmaf(x2.y, col[9], a2); a3 = fmaf(x3.y, col[13], a3);
    a0 = fmaf(x0.z, col[2], a0); a1 = fmaf(x1.z, col[6], a1); a2 = fmaf(x2.z, col[10], a2); a3 = fmaf(x3.z, col[14], a3);
    a0 = fmaf(x0.w, col[3], a0); a1 = fmaf(x1.w, col[7], a1); a2 = fmaf(x2.w, col[11], a2); a3 = fmaf(x3.w, col[15], a3);
    v = ((a0 + a1) + (a2 + a3)) * 0.9f;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// same step with 16 SHFL.IDX instead of shared memory
__global__ void k_shfl_step(float* out, long long* cyc) {
  const int lane = threadIdx.x & 31;
  float col[16]; for (int i = 0; i < 16; ++i) col[i] = 0.0625f + 1e-3f * ((lane + i) & 7);
  float v = 1.f + lane * 0.01f;
  long long t0 = clock64();
#pragma unroll 2
  for (int i = 0; i < N; ++i) {
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      a0 = fmaf(__shfl_sync(0xffffffffu, v, j, 16), col[j], a0);
      a1 = fmaf(__shfl_sync(0xffffffffu, v, j + 1, 16), col[j + 1], a1);
      a2 = fmaf(__shfl_sync(0xffffffffu, v, j + 2, 16), col[j + 2], a2);
      a3 = fmaf(__shfl_sync(0xffffffffu, v, j + 3, 16), col[j + 3], a3);
    }
    v = ((a0 + a1) + (a2 + a3)) * 0.9f;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_lds_lat(float* out, long long* cyc) {            // dependent LDS (pointer chase)
  __shared__ int nxt[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) nxt[i] = (i * 7 + 3) & 1023;
  __syncthreads();
  int p = threadIdx.x & 1023;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) p = nxt[p];
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = p;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_lds_tput(float* out, long long* cyc) {           // independent conflict-free LDS.32
  __shared__ float buf[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) buf[i] = i;
  __syncthreads();
  float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  const int l = threadIdx.x & 31;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N / 16; ++i) {
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      a0 += buf[(i * 16 + j) * 8 % 4064 + l]; a1 += buf[(i * 16 + j + 1) * 8 % 4064 + l];
      a2 += buf[(i * 16 + j + 2) * 8 % 4064 + l]; a3 += buf[(i * 16 + j + 3) * 8 % 4064 + l];
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  void* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024 * 8);
  long long h;
  const int nts[] = {32, 128, 256, 512, 1024};
#define RUN(K, T, label) for (int w = 0; w < 5; ++w) { K<<<1, nts[w]>>>((T*)out, cyc); K<<<1, nts[w]>>>((T*)out, cyc); cudaDeviceSynchronize(); \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("%-16s threads=%4d  cycles/iter(per warp)=%.2f  warp-instr/cyc/SM=%.3f\n", label, nts[w], (double)h / N, (double)N * (nts[w] / 32) / h); }
  RUN(k_shfl_tput, float, "shfl_indep");
  RUN(k_shfl_lat, float, "shfl_dep(+fadd)");
  RUN(k_dfma, double, "dfma_indep8");
  RUN(k_ffma, float, "ffma_indep8");
  RUN(k_ffma_chain, float, "ffma_chain");
  RUN(k_lds_lat, float, "lds_chase");
  RUN(k_lds_tput, float, "lds_indep");
  for (int w = 0; w < 3; ++w) { k_sts_lds_step<<<1, nts[w]>>>((float*)out, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-16s threads=%4d  cycles/step=%.2f\n", "sts_lds_step", nts[w], (double)h / N); }
  for (int w = 0; w < 3; ++w) { k_shfl_step<<<1, nts[w]>>>((float*)out, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-16s threads=%4d  cycles/step=%.2f\n", "shfl_step", nts[w], (double)h / N); }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
