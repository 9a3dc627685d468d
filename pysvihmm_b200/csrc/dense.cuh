// Dense K x K recursion step on the 5th-generation tensor cores (BASELINE config 4: K = 256, bf16):
// for 64 < K <= 256 the per-step matvec of 128 windows is one (128 x K) . (K x K) contraction, issued
// as tcgen05.mma (kind::f16, bf16 inputs, float32 accumulators in TENSOR MEMORY) by one elected thread
// per CTA.  Replaces, for the windows of a CTA, forward_msgs (hmmsgd_metaobs.py:775-803) and
// backward_msgs (:828-855); opt-in through SVIHMM_BF16_DENSE because the messages are rounded to
// bf16 between steps (marginals agree with the float64 reference to ~1e-2, not 1e-5).
//
//   grid (ceil(B/128), 2): CTA = 128 windows x one direction, 128 threads, thread = window = TMEM lane
//   shared memory  B operand: the transition matrix (forward: P^T, backward: P) as bf16, K-major,
//                  SWIZZLE_128B, resident for the whole kernel (KP x KP x 2 bytes, 128 KB at K = 256)
//                  A operand: the carried vectors of the 128 windows (128 x KP bf16), rewritten by
//                  the epilogue of every step in the same canonical layout
//   per step       thread 0: KP/16 x tcgen05.mma (M = 128, N = KP, K = 16) -> tcgen05.commit -> mbarrier
//                  all threads: tcgen05.ld their accumulator row (32 columns at a time), multiply by
//                  b[t] (and by an exact power of two taken from the previous step's row sum: the row
//                  sum is thread-local, no shuffles), write the float32 message to HBM and its bf16
//                  rounding into the A operand
// The descriptor encodings (shared-memory matrix descriptor, instruction descriptor) and the TMEM
// addressing were validated stand-alone by scripts/probes/umma_probe.cu (max error 9e-6 on a
// 128 x 256 x 256 bf16 GEMM against the host).
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

#define DN_M 128

__device__ __forceinline__ uint32_t dn_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// shared-memory matrix descriptor: K-major operand, 64-column (128-byte) slabs, SWIZZLE_128B
__device__ __forceinline__ uint64_t dn_desc(const uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);   // start address >> 4                       bits [0,14)
  d |= (uint64_t)1 << 16;                   // leading byte offset (unused: swizzled K-major)  [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;         // stride byte offset: 8 rows x 128 B        [32,46)
  d |= (uint64_t)1 << 46;                   // descriptor version 1 (sm_100)             [46,48)
  d |= (uint64_t)2 << 61;                   // SWIZZLE_128B                              [61,64)
  return d;
}
// byte offset of the 16-byte chunk holding columns [8c, 8c+8) of row r (operand with `rows` rows)
__device__ __forceinline__ uint32_t dn_chunk(const int r, const int c, const int rows) {
  return (uint32_t)(c >> 3) * rows * 128 + (r >> 3) * 1024 + (r & 7) * 128 + ((((c & 7) ^ (r & 7)) << 4));
}
__device__ __forceinline__ uint32_t dn_pack(const float a, const float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}

// Mrow: the B operand source, row-major [n][k] float (forward: P^T, backward: P).  Tables as in
// k_chain_wide: alpha_out scaled by 2^-E (E per row), beta_out arbitrary power-of-two scale per row.
__global__ void __launch_bounds__(DN_M)
k_chain_dense(int B, int T, int K, int KP, const float* __restrict__ Pfwd, const float* __restrict__ Pbwd,
              const float* __restrict__ pi0, const float* __restrict__ b, float* __restrict__ alpha_out,
              float* __restrict__ beta_out, int* __restrict__ E_out) {
  extern __shared__ __align__(1024) uint8_t dsm_raw[];
  uint8_t* dsm = dsm_raw + ((1024u - (dn_smem(dsm_raw) & 1023u)) & 1023u);   // SWIZZLE_128B atoms: 1024-byte aligned
  uint8_t* sA = dsm;                               // DN_M x KP bf16
  uint8_t* sB = dsm + (size_t)DN_M * KP * 2;       // KP x KP bf16
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, wp = tid >> 5;
  const bool fwd = blockIdx.y == 0;
  const int w = blockIdx.x * DN_M + tid;            // this thread's window
  const bool live = w < B;
  const float* Mrow = fwd ? Pfwd : Pbwd;
  // B operand: row n, column k (zero padded to KP x KP)
  for (int i = tid; i < KP * (KP / 8); i += DN_M) {
    const int n = i / (KP / 8), c = i - n * (KP / 8);
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { const int k = 8 * c + u; v[u] = (n < K && k < K) ? __ldg(Mrow + (size_t)n * K + k) : 0.f; }
    *reinterpret_cast<uint4*>(sB + dn_chunk(n, c, KP)) = make_uint4(dn_pack(v[0], v[1]), dn_pack(v[2], v[3]), dn_pack(v[4], v[5]), dn_pack(v[6], v[7]));
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dn_smem(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (wp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dn_smem(&tmem_base)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // step 0: forward alpha_0 = pi0 * b_0; backward beta_{T-1} = 1, carried b_{T-1}
  const int dt = fwd ? 1 : -1;
  int t = fwd ? 0 : T - 1;
  const size_t wbase = (size_t)(live ? w : 0) * T * K;
  float* outp = (fwd ? alpha_out : beta_out) + wbase;
  const float* bw = b + wbase;
  float sum = 0.f;
  for (int c = 0; c < KP / 8; ++c) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = 8 * c + u;
      const float bb = (live && k < K) ? bw[(size_t)t * K + k] : 0.f;
      v[u] = fwd ? ((live && k < K) ? __ldg(pi0 + k) * bb : 0.f) : bb;
      if (live && k < K) outp[(size_t)t * K + k] = fwd ? v[u] : 1.f;
      sum += v[u];
    }
    *reinterpret_cast<uint4*>(sA + dn_chunk(tid, c, DN_M)) = make_uint4(dn_pack(v[0], v[1]), dn_pack(v[2], v[3]), dn_pack(v[4], v[5]), dn_pack(v[6], v[7]));
  }
  int E = 0;
  if (fwd && live) E_out[(size_t)w * T + t] = 0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(KP >> 3) << 17) | ((uint32_t)(DN_M >> 4) << 24);
  const uint32_t aA = dn_smem(sA), aB = dn_smem(sB);
  uint32_t phase = 0;
  for (int s = 1; s < T; ++s) {
    t += dt;
    if (tid == 0) {
      for (int ks = 0; ks < KP / 16; ++ks) {
        const uint64_t da = dn_desc(aA + (ks >> 2) * (DN_M * 128) + (ks & 3) * 32);
        const uint64_t db = dn_desc(aB + (ks >> 2) * (KP * 128) + (ks & 3) * 32);
        const uint32_t acc = ks > 0;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(dn_smem(&bar)) : "memory");
    }
    // exact power-of-two rescale from the previous step's row sum (thread-local)
    int d = (int)((__float_as_uint(sum) >> 23) & 0xff) - 127;
    if (!(sum > 0.f)) d = 0;
    d = max(-100, min(100, d));
    const float r = __uint_as_float((unsigned)(127 - d) << 23);
    E += d;
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(dn_smem(&bar)), "r"(phase) : "memory");
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    sum = 0.f;
    // b of this step was pulled into L2 two steps ago; pull the row of step s+2 now (8 x 128 B per thread)
    if (live && s + 2 < T) {
      const char* pf = reinterpret_cast<const char*>(bw + (size_t)(t + 2 * dt) * K);
      for (int o = 0; o < K * 4; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + o));
    }
    float4 bq[8];                                      // b of the chunk being processed, fetched one chunk ahead
#pragma unroll
    for (int q4 = 0; q4 < 8; ++q4) {
      const int k = q4 * 4;
      bq[q4] = (live && k < K) ? *reinterpret_cast<const float4*>(bw + (size_t)t * K + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int c32 = 0; c32 < KP / 32; ++c32) {
      uint32_t a[32];
      const uint32_t ta = tm + ((uint32_t)(wp * 32) << 16) + c32 * 32;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]), "=r"(a[9]),
                     "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15]), "=r"(a[16]), "=r"(a[17]), "=r"(a[18]), "=r"(a[19]),
                     "=r"(a[20]), "=r"(a[21]), "=r"(a[22]), "=r"(a[23]), "=r"(a[24]), "=r"(a[25]), "=r"(a[26]), "=r"(a[27]), "=r"(a[28]), "=r"(a[29]),
                     "=r"(a[30]), "=r"(a[31]) : "r"(ta) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float4 bn[8];                                    // next chunk's b: in flight while this chunk is processed
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) {
        const int k = (c32 + 1) * 32 + q4 * 4;
        bn[q4] = (live && k < K) ? *reinterpret_cast<const float4*>(bw + (size_t)t * K + k) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) {                 // 4 columns at a time: one STG.128
        const int k = c32 * 32 + q4 * 4;
        const float4 bb = bq[q4];
        const float m0 = __uint_as_float(a[q4 * 4]) * r, m1 = __uint_as_float(a[q4 * 4 + 1]) * r;
        const float m2 = __uint_as_float(a[q4 * 4 + 2]) * r, m3 = __uint_as_float(a[q4 * 4 + 3]) * r;
        const float v0 = m0 * bb.x, v1 = m1 * bb.y, v2 = m2 * bb.z, v3 = m3 * bb.w;
        if (live && k < K) *reinterpret_cast<float4*>(outp + (size_t)t * K + k) = fwd ? make_float4(v0, v1, v2, v3) : make_float4(m0, m1, m2, m3);
        sum += (v0 + v1) + (v2 + v3);
        a[q4 * 4] = __float_as_uint(v0); a[q4 * 4 + 1] = __float_as_uint(v1);
        a[q4 * 4 + 2] = __float_as_uint(v2); a[q4 * 4 + 3] = __float_as_uint(v3);
      }
#pragma unroll
      for (int q4 = 0; q4 < 8; ++q4) bq[q4] = bn[q4];
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8)                   // the carried vector, rounded to bf16, back into the A operand
        *reinterpret_cast<uint4*>(sA + dn_chunk(tid, c32 * 4 + c8, DN_M)) =
            make_uint4(dn_pack(__uint_as_float(a[c8 * 8]), __uint_as_float(a[c8 * 8 + 1])),
                       dn_pack(__uint_as_float(a[c8 * 8 + 2]), __uint_as_float(a[c8 * 8 + 3])),
                       dn_pack(__uint_as_float(a[c8 * 8 + 4]), __uint_as_float(a[c8 * 8 + 5])),
                       dn_pack(__uint_as_float(a[c8 * 8 + 6]), __uint_as_float(a[c8 * 8 + 7])));
    }
    if (fwd && live) E_out[(size_t)w * T + t] = E;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (wp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256) : "memory");
}

// q[r] = alpha[r]*beta[r] / sum, lt[r] = log(sum_k alpha[r][k]) + E[r] ln 2 for any K (warp per row)
__global__ void __launch_bounds__(256)
k_marginals_any(int64_t R, int K, const float* __restrict__ alpha, const float* __restrict__ beta,
                const int* __restrict__ E, float* __restrict__ q, double* __restrict__ lt) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = w0; r < R; r += nw) {
    float sa = 0.f, sp = 0.f;
    for (int k = lane; k < K; k += 32) { const float al = alpha[r * K + k]; sa += al; sp += al * beta[r * K + k]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sa += __shfl_xor_sync(0xffffffffu, sa, o); sp += __shfl_xor_sync(0xffffffffu, sp, o); }
    const float inv = 1.f / sp;
    for (int k = lane; k < K; k += 32) q[r * K + k] = alpha[r * K + k] * beta[r * K + k] * inv;
    if (lane == 0) lt[r] = (double)logf(sa) + (double)E[r] * M_LN2;
  }
}
