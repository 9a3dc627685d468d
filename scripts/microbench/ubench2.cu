// Microbenchmarks, part 2: mma.sync tf32 latency/throughput, packed FFMA2, STS->LDS round trip,
// LDS.128 broadcast throughput, mixed shuffle+smem broadcast step.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096

__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__global__ void k_mma_chain(float* out, long long* cyc) {      // dependent accumulator chain
  float d[4] = {0, 0, 0, 0}; unsigned a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1e-3f * (threadIdx.x + i));
  b[0] = __float_as_uint(0.5f); b[1] = __float_as_uint(0.25f);
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) mma_tf32(d, a, b);
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = d[0] + d[1] + d[2] + d[3];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_mma_dep_a(float* out, long long* cyc) {      // output feeds next A operand
  float d[4] = {0, 0, 0, 0}; unsigned a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1e-3f * (threadIdx.x + i));
  b[0] = __float_as_uint(0.5f); b[1] = __float_as_uint(0.25f);
  long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) {
    d[0] = d[1] = d[2] = d[3] = 0.f;
    mma_tf32(d, a, b);
    a[0] = __float_as_uint(d[0]); a[1] = __float_as_uint(d[2]); a[2] = __float_as_uint(d[1]); a[3] = __float_as_uint(d[3]);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = d[0] + d[1] + d[2] + d[3];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_mma_indep(float* out, long long* cyc) {      // 8 independent accumulators
  float d[8][4]; unsigned a[4], b[2];
  for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) d[j][i] = 0;
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1e-3f * (threadIdx.x + i));
  b[0] = __float_as_uint(0.5f); b[1] = __float_as_uint(0.25f);
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N / 8; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) mma_tf32(d[j], a, b);
  }
  long long t1 = clock64();
  float s = 0; for (int j = 0; j < 8; ++j) s += d[j][0] + d[j][1] + d[j][2] + d[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_ffma2(float* out, long long* cyc) {          // 8 independent packed chains
  float2 a[8]; for (int j = 0; j < 8; ++j) a[j] = make_float2(threadIdx.x + j, j);
  const float2 m = make_float2(1.0000001f, 0.999999f), c = make_float2(1e-9f, 1e-8f);
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N / 8; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      unsigned long long& A = reinterpret_cast<unsigned long long&>(a[j]);
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A)
                   : "l"(reinterpret_cast<const unsigned long long&>(m)), "l"(reinterpret_cast<const unsigned long long&>(c)));
    }
  }
  long long t1 = clock64();
  float s = 0; for (int j = 0; j < 8; ++j) s += a[j].x + a[j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_ffma2_chain(float* out, long long* cyc) {
  float2 a = make_float2(threadIdx.x, 1.f);
  const float2 m = make_float2(1.0000001f, 0.999999f), c = make_float2(1e-9f, 1e-8f);
  unsigned long long& A = reinterpret_cast<unsigned long long&>(a);
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i)
    asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(A)
                 : "l"(reinterpret_cast<const unsigned long long&>(m)), "l"(reinterpret_cast<const unsigned long long&>(c)));
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a.x + a.y;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_sts_lds_rt(float* out, long long* cyc) {     // minimal STS -> LDS.32 (other lane) round trip
  __shared__ float buf[2][32][32];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  float v = lane;
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) {
    float* b = buf[i & 1][wp];
    b[lane] = v;
    __syncwarp();
    v = b[lane ^ 1] + 1.f;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_lds128_bcast(float* out, long long* cyc) {   // independent LDS.128, 2 distinct addresses per warp
  __shared__ __align__(16) float buf[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) buf[i] = i;
  __syncthreads();
  const int g = (threadIdx.x & 31) >> 4;
  float4 acc = make_float4(0, 0, 0, 0);
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N / 16; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4 x = *reinterpret_cast<const float4*>(buf + ((j * 32 + g * 16 + (i & 3) * 4) & 1023));
      acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_lds32_cf(float* out, long long* cyc) {       // independent conflict-free LDS.32
  __shared__ float buf[2048];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = i;
  __syncthreads();
  const int l = threadIdx.x & 31;
  float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N / 16; ++i) {
    const float* p = buf + l + (i & 15) * 32;
#pragma unroll
    for (int j = 0; j < 16; j += 4) { a0 += p[j * 32]; a1 += p[(j + 1) * 32]; a2 += p[(j + 2) * 32]; a3 += p[(j + 3) * 32]; }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// smem broadcast step with packed FFMA2 and the group max on the ALU pipe (candidate phase-B step)
__global__ void k_step_ffma2(float* out, long long* cyc) {
  __shared__ __align__(16) float buf[2][8][32];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5, g = lane >> 4;
  float2 col[8]; for (int i = 0; i < 8; ++i) col[i] = make_float2(0.0625f + 1e-3f * ((lane + i) & 7), 0.06f);
  float v = 1.f + lane * 0.01f; unsigned mxacc = 0;
  long long t0 = clock64();
#pragma unroll 2
  for (int i = 0; i < N; ++i) {
    float* b = buf[i & 1][wp];
    b[lane] = v;
    __syncwarp();
    const float4* p = reinterpret_cast<const float4*>(b + 16 * g);
    float4 x[4]; x[0] = p[0]; x[1] = p[1]; x[2] = p[2]; x[3] = p[3];
    unsigned long long acc0 = 0ull, acc1 = 0ull;
    unsigned m = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const unsigned long long lo = reinterpret_cast<const unsigned long long*>(&x[q])[0];
      const unsigned long long hi = reinterpret_cast<const unsigned long long*>(&x[q])[1];
      asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc0) : "l"(lo), "l"(reinterpret_cast<const unsigned long long&>(col[2 * q])));
      asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc1) : "l"(hi), "l"(reinterpret_cast<const unsigned long long&>(col[2 * q + 1])));
      m = max(max(m, __float_as_uint(x[q].x)), max(__float_as_uint(x[q].y), max(__float_as_uint(x[q].z), __float_as_uint(x[q].w))));
    }
    const float2 s0 = reinterpret_cast<const float2&>(acc0), s1 = reinterpret_cast<const float2&>(acc1);
    mxacc += m >> 23;
    v = ((s0.x + s0.y) + (s1.x + s1.y)) * 0.9f;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = v + mxacc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
// half of the vector by SHFL, half by shared memory
__global__ void k_step_mixed(float* out, long long* cyc) {
  __shared__ __align__(16) float buf[2][8][32];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5, g = lane >> 4;
  float col[16]; for (int i = 0; i < 16; ++i) col[i] = 0.0625f + 1e-3f * ((lane + i) & 7);
  float v = 1.f + lane * 0.01f;
  long long t0 = clock64();
#pragma unroll 2
  for (int i = 0; i < N; ++i) {
    float* b = buf[i & 1][wp];
    b[lane] = v;
    __syncwarp();
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      a0 = fmaf(__shfl_sync(0xffffffffu, v, j, 16), col[j], a0);
      a1 = fmaf(__shfl_sync(0xffffffffu, v, j + 1, 16), col[j + 1], a1);
    }
    const float4* p = reinterpret_cast<const float4*>(b + 16 * g + 8);
    const float4 x2 = p[0], x3 = p[1];
    a2 = fmaf(x2.x, col[8], a2); a3 = fmaf(x3.x, col[12], a3);
    a2 = fmaf(x2.y, col[9], a2); a3 = fmaf(x3.y, col[13], a3);
    a2 = fmaf(x2.z, col[10], a2); a3 = fmaf(x3.z, col[14], a3);
    a2 = fmaf(x2.w, col[11], a2); a3 = fmaf(x3.w, col[15], a3);
    v = ((a0 + a1) + (a2 + a3)) * 0.9f;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  void* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024 * 8);
  long long h;
  const int nts[] = {32, 128, 256, 512, 1024};
#define RUN(K, T, label, NW) for (int w = 0; w < NW; ++w) { K<<<1, nts[w]>>>((T*)out, cyc); K<<<1, nts[w]>>>((T*)out, cyc); cudaDeviceSynchronize(); \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("%-16s threads=%4d  cycles/op(per warp)=%.2f  warp-ops/cyc/SM=%.3f\n", label, nts[w], (double)h / N, (double)N * (nts[w] / 32) / h); }
  RUN(k_mma_chain, float, "mma_tf32_chainC", 5);
  RUN(k_mma_dep_a, float, "mma_tf32_depA", 2);
  RUN(k_mma_indep, float, "mma_tf32_indep8", 5);
  RUN(k_ffma2, float, "ffma2_indep8", 5);
  RUN(k_ffma2_chain, float, "ffma2_chain", 2);
  RUN(k_sts_lds_rt, float, "sts_lds_rt", 3);
  RUN(k_lds128_bcast, float, "lds128_bcast", 5);
  RUN(k_lds32_cf, float, "lds32_cf", 5);
  RUN(k_step_ffma2, float, "step_ffma2", 3);
  RUN(k_step_mixed, float, "step_mixed", 3);
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
