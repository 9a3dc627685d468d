// Probe: gather 256 windows of 16 KB out of a page-locked, mapped host series into a dense device buffer:
// (a) the zero-copy load/store kernel of svihmm_prefetch_windows (8 CTAs x 256 threads, 8 x 16-byte loads in
// flight per thread), (b) the bulk-copy engine: one thread per CTA issues cp.async.bulk global(host) ->
// shared per piece, then cp.async.bulk shared -> global; a ring of slots keeps pieces in flight.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_gather_probe tma_gather_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); } } while (0)

__global__ void __launch_bounds__(256) k_gather(int B, size_t wbytes, const uint8_t* __restrict__ src, const int64_t* __restrict__ off,
                                                uint8_t* __restrict__ dst) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  const size_t upw = wbytes / 16, total = upw * B;
  for (size_t i0 = tid; i0 < total; i0 += nth * 8) {
    int4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { const size_t i = i0 + (size_t)u * nth; if (i < total) { const size_t b = i / upw, k = i - b * upw; v[u] = *((const int4*)(src + off[b]) + k); } }
#pragma unroll
    for (int u = 0; u < 8; ++u) { const size_t i = i0 + (size_t)u * nth; if (i < total) ((int4*)dst)[i] = v[u]; }
  }
}

__device__ __forceinline__ uint32_t smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// piece = `pb` bytes; CTA c takes pieces c, c + grid, ...; ring of `ns` slots
__global__ void __launch_bounds__(32) k_bulk(int B, int wbytes, int pb, int ns, const uint8_t* __restrict__ src, const int64_t* __restrict__ off,
                                             uint8_t* __restrict__ dst) {
  extern __shared__ __align__(128) uint8_t ring[];
  __shared__ __align__(8) unsigned long long full[16];
  if (threadIdx.x != 0) return;
  for (int s = 0; s < ns; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem(full + s)) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const int ppw = wbytes / pb, np = B * ppw;
  int issued = 0, done = 0;
  const int mine = np > (int)blockIdx.x ? (np - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  while (done < mine) {
    // keep the ring full: slot s may be refilled once the store that read it has finished reading
    while (issued < mine && issued < done + ns) {
      if (issued >= ns) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(0) : "memory");   // simple: all prior stores have read their slots
      const int p = blockIdx.x + issued * gridDim.x, b = p / ppw, k = p - b * ppw, s = issued % ns;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem(full + s)), "r"(pb) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(smem(ring + (size_t)s * pb)), "l"(src + off[b] + (size_t)k * pb), "r"(pb), "r"(smem(full + s)) : "memory");
      ++issued;
    }
    const int s = done % ns; const unsigned par = (done / ns) & 1;
    unsigned ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem(full + s)), "r"(par) : "memory");
    const int p = blockIdx.x + done * gridDim.x;
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + (size_t)p * pb), "r"(smem(ring + (size_t)s * pb)), "r"(pb) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    ++done;
  }
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(0) : "memory");
}

int main() {
  const int B = 256; const size_t W = 16384, TOT = (size_t)1 << 28;
  cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  uint8_t* dev; CK(cudaMalloc(&dev, B * W)); uint8_t* ref = (uint8_t*)malloc(B * W); uint8_t* got = (uint8_t*)malloc(B * W);
  uint8_t* h = (uint8_t*)malloc(TOT); for (size_t i = 0; i < TOT; i += 4) *(uint32_t*)(h + i) = (uint32_t)(i * 2654435761u);
  CK(cudaHostRegister(h, TOT, cudaHostRegisterMapped)); uint8_t* hd; CK(cudaHostGetDevicePointer((void**)&hd, h, 0));
  std::vector<int64_t> off(B); for (int i = 0; i < B; ++i) { off[i] = (int64_t)(((size_t)i * 7919 * 4096 + 32 * (i % 5)) % (TOT - W)); memcpy(ref + i * W, h + off[i], W); }
  int64_t* doff; CK(cudaMalloc(&doff, B * 8)); CK(cudaMemcpy(doff, off.data(), B * 8, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](const char* label, auto launch) {
    CK(cudaMemset(dev, 0, B * W)); launch(); CK(cudaStreamSynchronize(st));
    CK(cudaMemcpy(got, dev, B * W, cudaMemcpyDeviceToHost)); const bool okc = memcmp(got, ref, B * W) == 0;
    float best = 1e9f;
    for (int r = 0; r < 10; ++r) { cudaEventRecord(e0, st); launch(); cudaEventRecord(e1, st); cudaStreamSynchronize(st); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    printf("%-52s %7.1f us  %5.1f GB/s  %s  %s\n", label, best * 1e3, B * W / (best * 1e-3) / 1e9, okc ? "ok" : "MISMATCH", cudaGetErrorString(cudaGetLastError()));
  };
  for (int ctas : {8, 16, 32}) { char l[96]; snprintf(l, 96, "zero-copy kernel, %d CTAs x 256 thr", ctas); timeit(l, [&] { k_gather<<<ctas, 256, 0, st>>>(B, W, hd, doff, dev); }); }
  CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int ctas : {4, 8, 16, 32})
    for (int pb : {4096, 16384})
      for (int ns : {2, 4, 8}) {
        if ((size_t)pb * ns > 200 * 1024) continue;
        char l[96]; snprintf(l, 96, "bulk engine, %d CTAs, %d B pieces, %d slots", ctas, pb, ns);
        timeit(l, [&] { k_bulk<<<ctas, 32, (size_t)pb * ns, st>>>(B, (int)W, pb, ns, hd, doff, dev); });
      }
  { float best = 1e9f; uint8_t* hp; CK(cudaMallocHost(&hp, B * W));
    for (int r = 0; r < 5; ++r) { cudaEventRecord(e0, st); cudaMemcpyAsync(dev, hp, B * W, cudaMemcpyHostToDevice, st); cudaEventRecord(e1, st); cudaStreamSynchronize(st); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    printf("%-52s %7.1f us  %5.1f GB/s\n", "one 4 MB cudaMemcpyAsync (pinned)", best * 1e3, B * W / (best * 1e-3) / 1e9); }
  return 0;
}
