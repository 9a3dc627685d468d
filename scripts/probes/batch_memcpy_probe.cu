// Probe: which argument combinations does cudaMemcpyBatchAsync accept for host->device slabs, and
// how fast is it against a zero-copy gather kernel?  (build: nvcc -arch=sm_100a -o probe this.cu)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <chrono>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); } } while (0)

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main() {
  const size_t N = 256, SZ = 16384, TOT = (size_t)1 << 28;
  cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  char* dev; CK(cudaMalloc(&dev, N * SZ));
  char* hreg = (char*)malloc(TOT); memset(hreg, 1, TOT);
  CK(cudaHostRegister(hreg, TOT, cudaHostRegisterMapped));
  char* hpin; CK(cudaMallocHost(&hpin, TOT)); memset(hpin, 2, TOT);
  int can = 0; CK(cudaDeviceGetAttribute(&can, cudaDevAttrCanUseHostPointerForRegisteredMem, 0));
  printf("canUseHostPointerForRegisteredMem=%d\n", can);
  std::vector<void*> d(N), s(N); std::vector<size_t> z(N, SZ);
  for (int variant = 0; variant < 6; ++variant) {
    char* base = (variant & 1) ? hpin : hreg;
    for (size_t i = 0; i < N; ++i) { d[i] = dev + i * SZ; s[i] = base + ((i * 7919 * 4096) % (TOT - SZ)); }
    cudaMemcpyAttributes at; memset(&at, 0, sizeof at);
    at.srcAccessOrder = (variant / 2) == 0 ? cudaMemcpySrcAccessOrderStream : ((variant / 2) == 1 ? cudaMemcpySrcAccessOrderAny : cudaMemcpySrcAccessOrderDuringApiCall);
    size_t idx = 0, fail = 999;
    cudaError_t e = cudaMemcpyBatchAsync(d.data(), s.data(), z.data(), N, &at, &idx, 1, &fail, st);
    printf("variant %d (src=%s order=%d): %s fail=%zu\n", variant, (variant & 1) ? "cudaMallocHost" : "registered", (int)at.srcAccessOrder, cudaGetErrorString(e), fail);
    cudaGetLastError();
    if (e == cudaSuccess) {
      CK(cudaStreamSynchronize(st));
      double best = 1e9, host = 0;
      for (int r = 0; r < 10; ++r) {
        double t0 = now();
        cudaMemcpyBatchAsync(d.data(), s.data(), z.data(), N, &at, &idx, 1, &fail, st);
        double t1 = now();
        cudaStreamSynchronize(st);
        double t2 = now();
        if (t2 - t0 < best) { best = t2 - t0; host = t1 - t0; }
      }
      printf("   %zu x %zu B: %.1f us total (%.1f GB/s), host submit %.1f us\n", N, SZ, best * 1e6, N * SZ / best / 1e9, host * 1e6);
    }
  }
  // plain cudaMemcpyAsync per slab for comparison
  { double t0 = now(); for (size_t i = 0; i < N; ++i) cudaMemcpyAsync(dev + i * SZ, hreg + ((i * 7919 * 4096) % (TOT - SZ)), SZ, cudaMemcpyHostToDevice, st);
    double t1 = now(); cudaStreamSynchronize(st); double t2 = now();
    printf("256 x cudaMemcpyAsync: %.1f us total, host submit %.1f us\n", (t2 - t0) * 1e6, (t1 - t0) * 1e6); }
  { double t0 = now(); cudaMemcpyAsync(dev, hpin, N * SZ, cudaMemcpyHostToDevice, st); cudaStreamSynchronize(st); double t2 = now();
    t0 = now(); cudaMemcpyAsync(dev, hpin, N * SZ, cudaMemcpyHostToDevice, st); cudaStreamSynchronize(st); t2 = now();
    printf("one 4 MB cudaMemcpyAsync (pinned): %.1f us (%.1f GB/s)\n", (t2 - t0) * 1e6, N * SZ / (t2 - t0) / 1e9); }
  return 0;
}
