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
//                  rounding into the A operand.  The step tables are kept window-minor ("tile
//                  layout", below) so that these per-thread accesses are one 128-byte line per warp.
// The descriptor encodings (shared-memory matrix descriptor, instruction descriptor) and the TMEM
// addressing were validated stand-alone by scripts/probes/umma_probe.cu (max error 9e-6 on a
// 128 x 256 x 256 bf16 GEMM against the host).
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

#define DN_M 128
#define DN_NS 4          // threads per window in k_chain_dense: each owns every DN_NS-th 32-column chunk

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

// Tile layout of the step tables (bT, alphaT, betaT): [tile][t][k][128 windows], so that thread = window
// accesses are fully coalesced (one 128-byte line per warp and state) and one step of a tile is one
// contiguous K x 512-byte block; ET is [tile][t][128].
__device__ __forceinline__ size_t dn_tile_off(const int tile, const int T, const int K, const int t) {
  return ((size_t)tile * T + t) * K * DN_M;
}

// b[w][t][k] (row-major, k_ll_to_b) -> bT in tile layout; rows of windows beyond B are zero
__global__ void __launch_bounds__(256)
k_dense_tile_b(int B, int T, int K, const float* __restrict__ b, float* __restrict__ bT) {
  __shared__ float s[32][DN_M + 1];
  const int t = blockIdx.x, tile = blockIdx.y, lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  float* out = bT + dn_tile_off(tile, T, K, t);
  for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
      const int wl = wp * 16 + i, w = tile * DN_M + wl, k = k0 + lane;
      s[lane][wl] = (w < B && k < K) ? __ldg(b + ((size_t)w * T + t) * K + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kk = wp * 4 + i;
      if (k0 + kk < K)
#pragma unroll
        for (int j = 0; j < 4; ++j) out[(size_t)(k0 + kk) * DN_M + j * 32 + lane] = s[kk][j * 32 + lane];
    }
    __syncthreads();
  }
}

// Mrow: the B operand source, row-major [n][k] float (forward: P^T, backward: P).  Tables in tile
// layout: alphaT scaled by 2^-E (E per row, ET), betaT arbitrary power-of-two scale per row.
__global__ void __launch_bounds__(DN_M * DN_NS)
k_chain_dense(int B, int T, int K, int KP, const float* __restrict__ Pfwd, const float* __restrict__ Pbwd,
              const float* __restrict__ pi0, const float* __restrict__ bT, float* __restrict__ alphaT,
              float* __restrict__ betaT, int* __restrict__ ET) {
  extern __shared__ __align__(1024) uint8_t dsm_raw[];
  uint8_t* dsm = dsm_raw + ((1024u - (dn_smem(dsm_raw) & 1023u)) & 1023u);   // SWIZZLE_128B atoms: 1024-byte aligned
  uint8_t* sA = dsm;                               // DN_M x KP bf16
  uint8_t* sB = dsm + (size_t)DN_M * KP * 2;       // KP x KP bf16
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base;
  __shared__ float psum[2][DN_NS][DN_M];           // partial row sums of the DN_NS threads of a window, by step parity
  const int tid = threadIdx.x & (DN_M - 1), wp = tid >> 5, part = threadIdx.x >> 7;
  const bool fwd = blockIdx.y == 0;
  const int tile = blockIdx.x;
  const float* Mrow = fwd ? Pfwd : Pbwd;
  // B operand: row n, column k (zero padded to KP x KP)
  for (int i = threadIdx.x; i < KP * (KP / 8); i += DN_M * DN_NS) {
    const int n = i / (KP / 8), c = i - n * (KP / 8);
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { const int k = 8 * c + u; v[u] = (n < K && k < K) ? __ldg(Mrow + (size_t)n * K + k) : 0.f; }
    *reinterpret_cast<uint4*>(sB + dn_chunk(n, c, KP)) = make_uint4(dn_pack(v[0], v[1]), dn_pack(v[2], v[3]), dn_pack(v[4], v[5]), dn_pack(v[6], v[7]));
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dn_smem(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dn_smem(&tmem_base)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // step 0: forward alpha_0 = pi0 * b_0; backward beta_{T-1} = 1, carried b_{T-1}
  const int dt = fwd ? 1 : -1;
  int t = fwd ? 0 : T - 1;
  const size_t tb = dn_tile_off(tile, T, K, 0) + tid;      // + (t*K + k)*DN_M
  float* outp = (fwd ? alphaT : betaT) + tb;
  const float* bw = bT + tb;
  int* Ep = ET + (size_t)tile * T * DN_M + tid;
  float sum = 0.f;
  if (part == 0)
  for (int c = 0; c < KP / 8; ++c) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = 8 * c + u;
      const float bb = (k < K) ? bw[((size_t)t * K + k) * DN_M] : 0.f;
      v[u] = fwd ? ((k < K) ? __ldg(pi0 + k) * bb : 0.f) : bb;
      if (k < K) outp[((size_t)t * K + k) * DN_M] = fwd ? v[u] : 1.f;
      sum += v[u];
    }
    *reinterpret_cast<uint4*>(sA + dn_chunk(tid, c, DN_M)) = make_uint4(dn_pack(v[0], v[1]), dn_pack(v[2], v[3]), dn_pack(v[4], v[5]), dn_pack(v[6], v[7]));
  }
  int E = 0;
  if (fwd && part == 0) Ep[(size_t)t * DN_M] = 0;
  psum[0][part][tid] = sum;
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
    sum = 0.f;
#pragma unroll
    for (int p2 = 0; p2 < DN_NS; ++p2) sum += psum[(s - 1) & 1][p2][tid];     // same order in every thread of the window
    if (threadIdx.x == 0) {
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
    const float* bt = bw + (size_t)t * K * DN_M;
    float* ot = outp + (size_t)t * K * DN_M;
    // the tile of step s+2 (contiguous K x 512 bytes) into L2: K*4 bytes per thread
    if (s + 2 < T) {
      const char* pf = reinterpret_cast<const char*>(bT + dn_tile_off(tile, T, K, t + 2 * dt)) + (size_t)threadIdx.x * K;
      for (int o = 0; o < K; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + o));
    }
    float bq[2][32];                                   // this thread's b of the whole step: in flight under the MMA
#pragma unroll
    for (int ci = 0; ci < 2; ++ci)
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int k = (part + ci * DN_NS) * 32 + j;
        bq[ci][j] = (k < K) ? bt[(size_t)k * DN_M] : 0.f;
      }
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(dn_smem(&bar)), "r"(phase) : "memory");
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int ci = 0; ci < 2; ++ci) {                   // KP <= 256: at most two 32-column chunks per thread
      const int c32 = part + ci * DN_NS;
      if (c32 >= KP / 32) break;
      uint32_t a[32];
      const uint32_t ta = tm + ((uint32_t)(wp * 32) << 16) + c32 * 32;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]), "=r"(a[9]),
                     "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15]), "=r"(a[16]), "=r"(a[17]), "=r"(a[18]), "=r"(a[19]),
                     "=r"(a[20]), "=r"(a[21]), "=r"(a[22]), "=r"(a[23]), "=r"(a[24]), "=r"(a[25]), "=r"(a[26]), "=r"(a[27]), "=r"(a[28]), "=r"(a[29]),
                     "=r"(a[30]), "=r"(a[31]) : "r"(ta) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const int k = c32 * 32 + j;
        const float m0 = __uint_as_float(a[j]) * r, m1 = __uint_as_float(a[j + 1]) * r;
        const float m2 = __uint_as_float(a[j + 2]) * r, m3 = __uint_as_float(a[j + 3]) * r;
        const float v0 = m0 * bq[ci][j], v1 = m1 * bq[ci][j + 1], v2 = m2 * bq[ci][j + 2], v3 = m3 * bq[ci][j + 3];
        if (k < K) {                                   // K % 4 == 0: whole groups of four
          ot[(size_t)k * DN_M] = fwd ? v0 : m0;
          ot[(size_t)(k + 1) * DN_M] = fwd ? v1 : m1;
          ot[(size_t)(k + 2) * DN_M] = fwd ? v2 : m2;
          ot[(size_t)(k + 3) * DN_M] = fwd ? v3 : m3;
        }
        s0 += v0; s1 += v1; s2 += v2; s3 += v3;
        a[j] = __float_as_uint(v0); a[j + 1] = __float_as_uint(v1);
        a[j + 2] = __float_as_uint(v2); a[j + 3] = __float_as_uint(v3);
      }
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8)                   // the carried vector, rounded to bf16, back into the A operand
        *reinterpret_cast<uint4*>(sA + dn_chunk(tid, c32 * 4 + c8, DN_M)) =
            make_uint4(dn_pack(__uint_as_float(a[c8 * 8]), __uint_as_float(a[c8 * 8 + 1])),
                       dn_pack(__uint_as_float(a[c8 * 8 + 2]), __uint_as_float(a[c8 * 8 + 3])),
                       dn_pack(__uint_as_float(a[c8 * 8 + 4]), __uint_as_float(a[c8 * 8 + 5])),
                       dn_pack(__uint_as_float(a[c8 * 8 + 6]), __uint_as_float(a[c8 * 8 + 7])));
    }
    if (fwd && part == 0) Ep[(size_t)t * DN_M] = E;
    psum[s & 1][part][tid] = (s0 + s1) + (s2 + s3);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256) : "memory");
}

// tile layout -> q[w][t][k] = alpha*beta / sum (row-major, coalesced through shared memory),
// lt[w][t] = log(sum_k alpha) + E ln 2, and optionally q16 = bf16(q) in tile layout.  One CTA per
// (t, tile), thread = window.
__global__ void __launch_bounds__(DN_M)
k_marginals_tiled(int B, int T, int K, const float* __restrict__ alphaT, const float* __restrict__ betaT,
                  const int* __restrict__ ET, float* __restrict__ q, double* __restrict__ lt,
                  __nv_bfloat16* __restrict__ q16) {
  __shared__ float s[32][DN_M + 1];
  const int t = blockIdx.x, tile = blockIdx.y, tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  const int w = tile * DN_M + tid;
  const size_t base = dn_tile_off(tile, T, K, t) + tid;
  const float* al = alphaT + base;
  const float* be = betaT + base;
  float sa = 0.f, sp = 0.f;
#pragma unroll 8
  for (int k = 0; k < K; ++k) { const float a = al[(size_t)k * DN_M]; sa += a; sp = fmaf(a, be[(size_t)k * DN_M], sp); }
  const float inv = 1.f / sp;
  if (w < B) lt[(size_t)w * T + t] = (double)logf(sa) + (double)ET[((size_t)tile * T + t) * DN_M + tid] * M_LN2;
  for (int k0 = 0; k0 < K; k0 += 32) {
#pragma unroll 8
    for (int kk = 0; kk < 32; ++kk) {
      const float v = (k0 + kk < K) ? al[(size_t)(k0 + kk) * DN_M] * be[(size_t)(k0 + kk) * DN_M] * inv : 0.f;
      s[kk][tid] = v;
      // the same marginal in tile layout, rounded to bf16: operand of k_tran_stats_dense (zero beyond B)
      if (q16 && k0 + kk < K) q16[base + (size_t)(k0 + kk) * DN_M] = __float2bfloat16_rn(w < B ? v : 0.f);
    }
    __syncthreads();
    for (int i = 0; i < 32; ++i) {
      const int wl = wp * 32 + i, ww = tile * DN_M + wl;
      if (ww < B && k0 + lane < K) q[((size_t)ww * T + t) * K + k0 + lane] = s[lane][wl];
    }
    __syncthreads();
  }
}

// Transition statistic of the dense path on the tensor cores (replaces the [0, K) columns of k_stats,
// i.e. hmmsgd_metaobs.py:876-878 with quirks Q1/Q2): A[i][j] = sum over windows w and pairs (t, t+1)
// (and (T-1, 0) when wrap) of q[w][t][i] q[w][t+1][j].  In tile layout the bf16 marginals of one
// (tile, t) are a K x 128 matrix with the reduction index (the window) contiguous, i.e. directly a
// K-major tcgen05 operand: one pair is Q_t (M = K rows, two halves of 128) times Q_{t+1} (N = K rows) over
// 128 windows = 16 tcgen05.mma of 128 x KP x 16, accumulated in TENSOR MEMORY across all pairs of the CTA
// (2 x KP float32 columns).  grid (tiles, TS): a CTA owns pairs [p0, p1) of one tile; three 64 KB
// operand slots (Q_t, Q_{t+1}, Q_{t+2} in flight by cp.async).  The epilogue stores the accumulators
// into the CTA's own split of the k_stats partials, summed in float64 and fixed order by k_stats_finalize.
#define TSD_SLOT 65536
__global__ void __launch_bounds__(256)
k_tran_stats_dense(int T, int K, int KP, int NP, int TS, const __nv_bfloat16* __restrict__ q16,
                   float* __restrict__ part, int N) {
  extern __shared__ __align__(1024) uint8_t tsm_raw[];
  uint8_t* sm = tsm_raw + ((1024u - (dn_smem(tsm_raw) & 1023u)) & 1023u);
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, wp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int p0 = (int)((int64_t)blockIdx.y * NP / TS), p1 = (int)((int64_t)(blockIdx.y + 1) * NP / TS);
  const int n = p1 - p0;
  if (n <= 0) return;
  for (int i = tid; i < 3 * TSD_SLOT / 16; i += 256) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dn_smem(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (wp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dn_smem(&tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;
  const uint32_t sbase = dn_smem(sm);
  auto load = [&](int t, int slot) {                  // Q_t: K rows x 128 windows (256 bytes per row) into a swizzled slot
    if (t >= T) t -= T;
    const __nv_bfloat16* src = q16 + dn_tile_off(tile, T, K, t);
    const uint32_t dst = sbase + (uint32_t)slot * TSD_SLOT;
    for (int i = tid; i < K * 16; i += 256) {
      const int r = i >> 4, c = i & 15;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + dn_chunk(r, c, 256)), "l"(src + (size_t)r * DN_M + c * 8) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load(p0, 0);
  load(p0 + 1, 1);
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(KP >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const int MH = (KP + 127) >> 7;
  uint32_t phase = 0;
  for (int i = 0; i < n; ++i) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (i >= 1) {                                      // the MMAs of pair i-1 are done: slot (i+2) % 3 is free again
      uint32_t ok = 0;
      while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(dn_smem(&bar)), "r"(phase) : "memory");
      phase ^= 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (i + 2 <= n) load(p0 + i + 2, (i + 2) % 3);
    if (tid == 0) {
      const uint32_t aA = sbase + (uint32_t)(i % 3) * TSD_SLOT, aB = sbase + (uint32_t)((i + 1) % 3) * TSD_SLOT;
      for (int h = 0; h < MH; ++h)
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t da = dn_desc(aA + (ks >> 2) * (256 * 128) + h * (128 * 128) + (ks & 3) * 32);
          const uint64_t db = dn_desc(aB + (ks >> 2) * (256 * 128) + (ks & 3) * 32);
          const uint32_t acc = (i > 0) || (ks > 0);
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tm + h * 256), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(dn_smem(&bar)) : "memory");
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(dn_smem(&bar)), "r"(phase) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // every CTA owns one split of the partials (plain stores; k_stats_finalize sums the splits in a fixed order in
  // float64: run-to-run deterministic, no float atomics)
  float* pout = part + (size_t)(blockIdx.x * gridDim.y + blockIdx.y) * K * N;
  // epilogue: warp (wp & 3) owns TMEM lanes 32 (wp & 3) .. +31; the two warp groups alternate 32-column chunks
  const int wq = wp & 3, wh = wp >> 2;
  for (int h = 0; h < MH; ++h) {
    const int row = h * 128 + wq * 32 + lane;
    for (int c32 = wh; c32 < KP / 32; c32 += 2) {
      uint32_t a[32];
      const uint32_t ta = tm + ((uint32_t)(wq * 32) << 16) + h * 256 + c32 * 32;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]), "=r"(a[9]),
                     "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15]), "=r"(a[16]), "=r"(a[17]), "=r"(a[18]), "=r"(a[19]),
                     "=r"(a[20]), "=r"(a[21]), "=r"(a[22]), "=r"(a[23]), "=r"(a[24]), "=r"(a[25]), "=r"(a[26]), "=r"(a[27]), "=r"(a[28]), "=r"(a[29]),
                     "=r"(a[30]), "=r"(a[31]) : "r"(ta) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row < K)
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = c32 * 32 + j;
          if (col < K) pout[(size_t)row * N + col] = __uint_as_float(a[j]);
        }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (wp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

// Emission statistics of the dense path on the tensor cores (replaces the [K, K + NB) columns of
// k_stats: n, sum w x, sum w x x^T / x^2; hmmsgd_metaobs.py:861-874 through the distribution's
// expected statistics): S[i][f] = sum over windows and t of q[w][t][i] F[w][t][f].  The features are
// built once per step in tile layout as bf16 hi + lo pairs (f = hi + lo to 2^-17, so only q carries
// bf16 rounding), one CTA-step is Q_t (K x 128 windows) times [F_hi ; F_lo] accumulated into the same
// TENSOR MEMORY columns.
#define ESD_NB 144                                    // feature rows of an operand slot (N of the MMA)
__global__ void __launch_bounds__(DN_M)
k_dense_tile_feat(int B, int T, int D, int NB, int diag, const void* __restrict__ obs, int dtype,
                  const uint8_t* __restrict__ mask, const int64_t* __restrict__ starts,
                  __nv_bfloat16* __restrict__ fhi, __nv_bfloat16* __restrict__ flo) {
  const int t = blockIdx.x, tile = blockIdx.y, tid = threadIdx.x;
  const int w = tile * DN_M + tid;
  const size_t base = ((size_t)tile * T + t) * NB * DN_M + tid;
  float wv = 0.f;
  int64_t r = 0;
  if (w < B) {
    r = starts[w] + t;
    wv = (mask && mask[r]) ? 0.f : 1.f;
    if (wv != 0.f) {
      bool bad = false;
      for (int d = 0; d < D; ++d) bad |= isnan((float)ld_obs(obs, dtype, r * D + d));
      if (bad) wv = 0.f;
    }
  }
  auto put = [&](const int f, const float v) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    fhi[base + (size_t)f * DN_M] = h;
    flo[base + (size_t)f * DN_M] = __float2bfloat16_rn(v - __bfloat162float(h));
  };
  put(0, wv);
  for (int d = 0; d < D; ++d) put(1 + d, wv != 0.f ? (float)ld_obs(obs, dtype, r * D + d) : 0.f);
  for (int m = 0; m < NB - 1 - D; ++m) {
    float v = 0.f;
    if (wv != 0.f) {
      if (diag) { const float x = (float)ld_obs(obs, dtype, r * D + m); v = x * x; }
      else { const int d1 = m / D, d2 = m - d1 * D; v = (float)ld_obs(obs, dtype, r * D + d1) * (float)ld_obs(obs, dtype, r * D + d2); }
    }
    put(1 + D + m, v);
  }
}

__global__ void __launch_bounds__(256)
k_emit_stats_dense(int T, int K, int KP, int NB, int TS, const __nv_bfloat16* __restrict__ q16,
                   const __nv_bfloat16* __restrict__ fhi, const __nv_bfloat16* __restrict__ flo,
                   float* __restrict__ part, int N, int col0) {
  extern __shared__ __align__(1024) uint8_t esm_raw[];
  uint8_t* sm = esm_raw + ((1024u - (dn_smem(esm_raw) & 1023u)) & 1023u);
  constexpr int BSLOT = ESD_NB * 256;                 // 144 rows x 128 windows bf16
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, wp = tid >> 5, lane = tid & 31;
  const int tile = blockIdx.x;
  const int t0 = (int)((int64_t)blockIdx.y * T / TS), t1 = (int)((int64_t)(blockIdx.y + 1) * T / TS);
  if (t1 <= t0) return;
  for (int i = tid; i < (TSD_SLOT + 2 * BSLOT) / 16; i += 256) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dn_smem(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (wp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dn_smem(&tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;
  const uint32_t aA = dn_smem(sm), aBh = aA + TSD_SLOT, aBl = aBh + BSLOT;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(ESD_NB >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const int MH = (KP + 127) >> 7;
  uint32_t phase = 0;
  for (int t = t0; t < t1; ++t) {
    const __nv_bfloat16* qs = q16 + dn_tile_off(tile, T, K, t);
    for (int i = tid; i < K * 16; i += 256) {
      const int r = i >> 4, c = i & 15;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(aA + dn_chunk(r, c, 256)), "l"(qs + (size_t)r * DN_M + c * 8) : "memory");
    }
    const size_t fo = ((size_t)tile * T + t) * NB * DN_M;
    for (int i = tid; i < NB * 16; i += 256) {
      const int r = i >> 4, c = i & 15;
      const uint32_t off = dn_chunk(r, c, ESD_NB);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(aBh + off), "l"(fhi + fo + (size_t)r * DN_M + c * 8) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(aBl + off), "l"(flo + fo + (size_t)r * DN_M + c * 8) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
      for (int h = 0; h < MH; ++h)
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t da = dn_desc(aA + (ks >> 2) * (256 * 128) + h * (128 * 128) + (ks & 3) * 32);
          const uint64_t dbh = dn_desc(aBh + (ks >> 2) * (ESD_NB * 128) + (ks & 3) * 32);
          const uint64_t dbl = dn_desc(aBl + (ks >> 2) * (ESD_NB * 128) + (ks & 3) * 32);
          const uint32_t acc = (t > t0) || (ks > 0);
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tm + h * 256), "l"(da), "l"(dbh), "r"(idesc), "r"(acc) : "memory");
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(tm + h * 256), "l"(da), "l"(dbl), "r"(idesc), "r"(1u) : "memory");
        }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(dn_smem(&bar)) : "memory");
    }
    uint32_t ok = 0;                                   // single stage: the slots are rewritten after the MMAs have read them
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(dn_smem(&bar)), "r"(phase) : "memory");
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  float* pout = part + (size_t)(blockIdx.x * gridDim.y + blockIdx.y) * K * N;     // this CTA's split (see k_tran_stats_dense)
  const int wq = wp & 3, wh = wp >> 2;
  for (int h = 0; h < MH; ++h) {
    const int row = h * 128 + wq * 32 + lane;
    for (int c32 = wh; c32 < (ESD_NB + 31) / 32; c32 += 2) {
      uint32_t a[32];
      const uint32_t ta = tm + ((uint32_t)(wq * 32) << 16) + h * 256 + c32 * 32;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                   : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]), "=r"(a[9]),
                     "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15]), "=r"(a[16]), "=r"(a[17]), "=r"(a[18]), "=r"(a[19]),
                     "=r"(a[20]), "=r"(a[21]), "=r"(a[22]), "=r"(a[23]), "=r"(a[24]), "=r"(a[25]), "=r"(a[26]), "=r"(a[27]), "=r"(a[28]), "=r"(a[29]),
                     "=r"(a[30]), "=r"(a[31]) : "r"(ta) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row < K)
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int col = c32 * 32 + j;
          if (col < NB) pout[(size_t)row * N + col0 + col] = __uint_as_float(a[j]);
        }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (wp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}
