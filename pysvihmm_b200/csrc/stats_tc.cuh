// Sufficient statistics of a minibatch on the 5th-generation tensor cores, operands fed by TMA
// (16 < K <= 64; BASELINE config 3): replaces k_stats_mma (Ampere-style mma.sync, 1.49 ms at c3).
//
//   S[f][k] = sum over the rows r of the minibatch of  F[r][f] * q[r][k]
//   F[r] = [ q[r+1] (K: the transition statistic, quirks Q1/Q2, hmmsgd_metaobs.py:873-881)
//          | w_r | w_r x_r (D) | w_r x_r,i x_r,j, i <= j  (util.py:73-83; w_r = 0 on masked / NaN rows) ]
// is ONE contraction over the rows.  A CTA walks tiles of 128 consecutive rows of one window:
//   TMA       cp.async.bulk.tensor.2d brings the 129 x K block of marginals (one halo row: the "next"
//             side of the last pair) and the 128 x D block of observations of the tile into shared
//             memory (tensor maps over the (B*T, K) marginals and over the (T_full, D) series; the
//             window start is a runtime coordinate), completion on mbarriers
//   operands  all 512 threads turn them into bf16 hi + lo pairs (v = hi + lo to 2^-17) in the K-major
//             SWIZZLE_128B layout of tcgen05: B operand = q^T (64 states x 128 rows), A operand = one
//             M-tile of 128 feature rows x 128 rows; two A buffers, so the generation of the next M-tile
//             overlaps the MMAs of the current one
//   MMA       one thread issues, per M-tile, 8 k-steps x 3 terms (lo.hi, hi.lo, hi.hi) of
//             tcgen05.mma.cta_group::1.kind::f16, M = 128, N = 64, float32 accumulators in TENSOR
//             MEMORY: nmt x 64 columns, kept across all tiles of the CTA; tcgen05.commit -> mbarriers
//   epilogue  tcgen05.ld of the accumulators, one partial [K][NF] per CTA; k_stats_sym_finalize sums the
//             partials in float64 (deterministic: no atomics anywhere)
// Only q and the features carry rounding: each to 2^-17 relative, independent across rows.
// Mixture emissions (BASELINE config 5): the B operand gets KE more columns, the weights q[r][k] r[r][k][c]
// of the K*C components (a third tensor map), so that one pass yields the transition statistic (feature
// rows "next q" x state columns) and the component statistics (feature rows [1 | x | xx] x component
// columns); tiles are 64 rows then, to fit shared memory.
#pragma once
#include <cuda.h>
#include "dense.cuh"

#define STC_NT 512            // 16 warps: the operand generation is issue-latency bound (ncu: 46 % of the issue slots with 8 warps)

struct StcArgs {
  int B, T, K, D, NF, diag, wrap, ntpw, nmt, ntiles;
  int KE;                    // mixture components (0: plain emissions)
  int RT;                    // rows per tile: 128, or 64 with mixture columns
  int N;                     // columns of the accumulators = MMA N: K (+ KE) rounded up to 16 (64 when plain)
  const float* q; const uint8_t* mask; const int64_t* starts;
  float* part;               // [gridDim.x][K + KE][NF]
};

struct StcSmem { size_t Bh, Bl, A, Asz, xs, qst, wst, wrow, qwrap, fa, fb, bars, total; };
__host__ __device__ inline StcSmem stc_layout(int K, int D, int KE, int RT, int N) {
  StcSmem s;
  const size_t bsz = (size_t)N * RT * 2;                // one B operand (hi or lo)
  s.Asz = (size_t)128 * RT * 2;                         // one A operand (hi or lo) of one buffer
  s.Bh = 0; s.Bl = bsz; s.A = 2 * bsz;                  // A: [buf][hi, lo]
  s.xs = s.A + 4 * s.Asz;
  s.qst = (s.xs + (size_t)RT * D * 4 + 127) & ~(size_t)127;
  s.wst = (s.qst + (size_t)(RT + 1) * K * 4 + 127) & ~(size_t)127;
  s.wrow = (s.wst + (size_t)RT * KE * 4 + 127) & ~(size_t)127;
  s.qwrap = s.wrow + 128 * 4;
  s.fa = s.qwrap + 64 * 4;
  s.fb = s.fa + 640;
  s.bars = (s.fb + 640 + 7) & ~(size_t)7;
  s.total = s.bars + 8 * 8 + 1024;                      // + slack for the 1024-byte alignment of the base
  return s;
}

__device__ __forceinline__ void stc_wait(unsigned long long* bar, const unsigned parity) {
  unsigned ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(dn_smem(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void stc_tma_2d(const CUtensorMap* tm, void* dst, unsigned long long* bar, const int c0, const int c1,
                                           const unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dn_smem(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dn_smem(dst)), "l"(tm), "r"(dn_smem(bar)), "r"(c0), "r"(c1) : "memory");
}
// 8 consecutive K-elements (rows of the minibatch) of one operand row -> bf16 hi and lo chunks
__device__ __forceinline__ void stc_pack(const float (&v)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * p], v[2 * p + 1]);
    const float2 hf = __bfloat1622float2(h2);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[2 * p] - hf.x, v[2 * p + 1] - hf.y);
    h[p] = *reinterpret_cast<const uint32_t*>(&h2); l[p] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]); lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(STC_NT, 1)
k_stats_tc(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_q,
           const __grid_constant__ CUtensorMap tm_w, const StcArgs a) {
  extern __shared__ __align__(1024) uint8_t stc_raw[];
  uint8_t* sm = stc_raw + ((1024u - (dn_smem(stc_raw) & 1023u)) & 1023u);
  const int K = a.K, D = a.D, T = a.T, NF = a.NF, KE = a.KE, RT = a.RT, N = a.N;
  const StcSmem L = stc_layout(K, D, KE, RT, N);
  uint8_t* sBh = sm + L.Bh; uint8_t* sBl = sm + L.Bl; uint8_t* sA = sm + L.A;
  float* xs = reinterpret_cast<float*>(sm + L.xs);
  float* qst = reinterpret_cast<float*>(sm + L.qst);
  float* wst = reinterpret_cast<float*>(sm + L.wst);
  float* wrow = reinterpret_cast<float*>(sm + L.wrow);
  float* qwrap = reinterpret_cast<float*>(sm + L.qwrap);
  uint8_t* fa = sm + L.fa; uint8_t* fb = sm + L.fb;
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + L.bars);
  unsigned long long* full_q = bars; unsigned long long* full_x = bars + 1;
  unsigned long long* mma_done = bars + 2;              // [2]: the MMAs that read A buffer 0 / 1
  unsigned long long* mma_all = bars + 4;               // all MMAs of a tile (B operand free again)
  unsigned long long* full_w = bars + 5;                // mixture weights of the tile
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  // feature table: fa = 255 "next q" column fb; 254 zero row; else value = w * xe[fa] * xe[fb], xe[D] = 1
  for (int f = tid; f < 640; f += STC_NT) {
    int ia = 254, ib = 0;
    if (f < K) { ia = 255; ib = f; }
    else if (f == K) { ia = D; ib = D; }
    else if (f < K + 1 + D) { ia = f - K - 1; ib = D; }
    else if (f < NF) {
      int e = f - (K + 1 + D);
      if (a.diag) { ia = e; ib = e; }
      else { int i = 0; while (e >= D - i) { e -= D - i; ++i; } ia = i; ib = i + e; }
    }
    fa[f] = (uint8_t)ia; fb[f] = (uint8_t)ib;
  }
  if (tid == 0) {
    for (int i = 0; i < 6; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dn_smem(bars + i)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_q) : "memory");
    if (KE) asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_w) : "memory");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dn_smem(&tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const unsigned qbytes = (unsigned)(RT + 1) * K * 4, xbytes = (unsigned)RT * D * 4, wbytes = (unsigned)RT * KE * 4;
  const int nchunk = RT / 8;                            // 16-byte chunks (8 rows) per operand row
  auto tile_w = [&](const int tau) { return tau / a.ntpw; };
  auto tile_t0 = [&](const int tau) { return (tau - (tau / a.ntpw) * a.ntpw) * RT; };
  int tau = blockIdx.x;
  if (tid == 0 && tau < a.ntiles) {
    const int w = tile_w(tau), t0 = tile_t0(tau);
    stc_tma_2d(&tm_q, qst, full_q, 0, w * T + t0, qbytes);
    stc_tma_2d(&tm_x, xs, full_x, 0, (int)(a.starts[w] + t0), xbytes);
    if (KE) stc_tma_2d(&tm_w, wst, full_w, 0, w * T + t0, wbytes);
  }
  unsigned it = 0, gen = 0;                             // tiles done by this CTA, A-operand generations so far
  for (; tau < a.ntiles; tau += gridDim.x, ++it) {
    const int w = tile_w(tau), t0 = tile_t0(tau);
    const int nrow = min(RT, T - t0);                   // real rows of the tile
    stc_wait(full_q, it & 1);
    if (KE) stc_wait(full_w, it & 1);
    if (it > 0) stc_wait(mma_all, (it - 1) & 1);        // the B operand is free again
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- B operand: [q | component weights]^T, hi and lo.  task = (column n, chunk of 8 rows)
    {
      for (int task = tid; task < N * nchunk; task += STC_NT) {
        const int n = task % N, c = task / N;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r = 8 * c + j;
          v[j] = r < nrow ? (n < K ? qst[r * K + n] : (n < K + KE ? wst[r * KE + (n - K)] : 0.f)) : 0.f;
        }
        uint4 hi, lo;
        stc_pack(v, hi, lo);
        const uint32_t off = dn_chunk(n, c, N);
        *reinterpret_cast<uint4*>(sBh + off) = hi; *reinterpret_cast<uint4*>(sBl + off) = lo;
      }
      // the wrap-around partner of the window's last row (quirk Q2): q[w][0]
      if (tid < 64) qwrap[tid] = (a.wrap && tid < K && t0 + RT >= T) ? __ldg(a.q + (size_t)w * T * K + tid) : 0.f;
    }
    // ---- row weights: 0 for masked rows, rows with a NaN, rows past the window end
    stc_wait(full_x, it & 1);
    {
      const int64_t g0 = a.starts[w] + t0;
      for (int r = wp * (RT / (STC_NT / 32)); r < (wp + 1) * (RT / (STC_NT / 32)); ++r) {
        const bool isn = lane < D ? isnan(xs[r * D + lane]) : false;
        const unsigned any = __ballot_sync(0xffffffffu, isn);
        if (lane == 0) wrow[r] = (r < nrow && !any && !(a.mask && a.mask[g0 + r])) ? 1.f : 0.f;
      }
    }
    __syncthreads();
    for (int m = 0; m < a.nmt; ++m, ++gen) {
      const unsigned ab = gen & 1, use = gen >> 1;
      if (use > 0) { stc_wait(mma_done + ab, (use - 1) & 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
      uint8_t* Ah = sA + (size_t)ab * 2 * L.Asz; uint8_t* Al = Ah + L.Asz;
      // ---- A operand of M-tile m: thread = (feature row fl, one part of the chunks of 8 rows)
      {
        constexpr int NPART = STC_NT / 128;
        const int fl = tid & 127, part = tid >> 7, f = 128 * m + fl;
        const int ka = fa[f], kb = fb[f];
#pragma unroll 2
        for (int c8 = 0; c8 < nchunk / NPART; ++c8) {
          const int c = part * (nchunk / NPART) + c8;
          float v[8];
          if (ka == 255) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int r = 8 * c + j, tn = t0 + r + 1;
              v[j] = tn < T ? qst[(r + 1) * K + kb] : (tn == T ? qwrap[kb] : 0.f);
            }
          } else if (ka == 254) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int r = 8 * c + j;
              const float xa = ka < D ? xs[r * D + ka] : 1.f, xb = kb < D ? xs[r * D + kb] : 1.f;
              v[j] = wrow[r] != 0.f ? xa * xb : 0.f;
            }
          }
          uint4 hi, lo;
          stc_pack(v, hi, lo);
          const uint32_t off = dn_chunk(fl, c, 128);
          *reinterpret_cast<uint4*>(Ah + off) = hi; *reinterpret_cast<uint4*>(Al + off) = lo;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (tid == 0) {
        const int taun = tau + gridDim.x;
        if (taun < a.ntiles) {
          // the staging buffers are free once their last readers are past the barrier above: the marginals
          // after M-tile 0 (the "next q" features), the observations after the last M-tile
          if (m == 0) {
            stc_tma_2d(&tm_q, qst, full_q, 0, tile_w(taun) * T + tile_t0(taun), qbytes);
            if (KE) stc_tma_2d(&tm_w, wst, full_w, 0, tile_w(taun) * T + tile_t0(taun), wbytes);
          }
          if (m == a.nmt - 1) stc_tma_2d(&tm_x, xs, full_x, 0, (int)(a.starts[tile_w(taun)] + tile_t0(taun)), xbytes);
        }
        const uint32_t aAh = dn_smem(Ah), aAl = dn_smem(Al), aBh = dn_smem(sBh), aBl = dn_smem(sBl);
        const uint32_t dcol = tm + (uint32_t)m * N;
#pragma unroll 1
        for (int ks = 0; ks < RT / 16; ++ks) {
          const uint32_t oa = (ks >> 2) * (128 * 128) + (ks & 3) * 32, ob = (ks >> 2) * (N * 128) + (ks & 3) * 32;
          const uint64_t dah = dn_desc(aAh + oa), dal = dn_desc(aAl + oa), dbh = dn_desc(aBh + ob), dbl = dn_desc(aBl + ob);
          const uint32_t acc0 = (it > 0 || ks > 0) ? 1u : 0u;
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(dcol), "l"(dal), "l"(dbh), "r"(idesc), "r"(acc0) : "memory");
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(dcol), "l"(dah), "l"(dbl), "r"(idesc), "r"(1u) : "memory");
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(dcol), "l"(dah), "l"(dbh), "r"(idesc), "r"(1u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(dn_smem(mma_done + ab)) : "memory");
        if (m == a.nmt - 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(dn_smem(mma_all)) : "memory");
      }
    }
  }
  // ---- epilogue: accumulators -> this CTA's partial [K + KE][NF]
  if (it > 0) stc_wait(mma_all, (it - 1) & 1);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    const int Kout = K + KE;
    float* part = a.part + (size_t)blockIdx.x * Kout * NF;
    const int quarter = wp & 3;
    for (int m = 0; m < a.nmt; ++m) {
      for (int cc = wp >> 2; cc < (N + 31) / 32; cc += STC_NT / 128) {       // 32-column chunks, STC_NT / 128 warps per lane quarter
        uint32_t v[32];
        const uint32_t ta = tm + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(m * N + cc * 32);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                       "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                       "=r"(v[30]), "=r"(v[31]) : "r"(ta) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int f = 128 * m + quarter * 32 + lane;
        if (f < NF) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int n = cc * 32 + j;
            if (n < Kout) part[(size_t)n * NF + f] = it > 0 ? __uint_as_float(v[j]) : 0.f;
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}

// partials of k_stats_tc with mixture columns, [z][K + KE][NF] -> packed statistics
// [ A (K*K) | n (KE) | sx (KE*D) | sxx (KE*DD) | q0 (K) | tail ]: the transition statistic from the state
// columns x "next q" feature rows, the component statistics from the component columns x emission feature
// rows (second moments stored as the upper triangle, mirrored here).  Float64 sums, fixed order.
__global__ void __launch_bounds__(256)
k_stats_tc_finalize_mix(int B, int T, int K, int KE, int D, int DD, int NF, int diag, int nsplit,
                        const float* __restrict__ part, const float* __restrict__ q,
                        const double* __restrict__ seq, const double* __restrict__ prior_tran, int add_prior,
                        double* __restrict__ out, size_t slen) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= slen) return;
  const int Kout = K + KE;
  const size_t o_n = (size_t)K * K, o_sx = o_n + KE, o_sxx = o_sx + (size_t)KE * D,
               o_q0 = o_sxx + (size_t)KE * DD, o_tail = o_q0 + K;
  int m = -1, n = 0;
  double v = 0.0;
  if (idx < o_n) { m = (int)(idx / K); n = (int)(idx % K); }
  else if (idx < o_sx) { m = K + (int)(idx - o_n); n = K; }
  else if (idx < o_sxx) { const size_t e = idx - o_sx; m = K + (int)(e / D); n = K + 1 + (int)(e % D); }
  else if (idx < o_q0) {
    const size_t e = idx - o_sxx; m = K + (int)(e / DD);
    const int c = (int)(e % DD);
    if (diag) n = K + 1 + D + c;
    else {
      int i = c / D, j = c - i * D;
      if (i > j) { const int tmp = i; i = j; j = tmp; }
      n = K + 1 + D + i * D - i * (i - 1) / 2 + (j - i);
    }
  }
  if (m >= 0) {
    for (int z = 0; z < nsplit; ++z) v += (double)part[((size_t)z * Kout + m) * NF + n];
    if (idx < o_n && add_prior) v += (double)B * (prior_tran[idx] - 1.0);
  } else if (idx < o_tail) {
    const int k = (int)(idx - o_q0);
    for (int b = 0; b < B; ++b) v += (double)q[(size_t)b * T * K + k];
  } else {
    const int tt = (int)(idx - o_tail);
    if (tt < 2) for (int b = 0; b < B; ++b) v += seq[2 * b + tt];
    else if (tt == 2) v = (double)B;
  }
  out[idx] = v;
}
