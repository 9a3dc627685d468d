// Per-phase kernels for windows that do not fit in shared memory with 32 < K <= 64 (BASELINE
// config 3: K = 64 full-covariance, D = 32, T = 1024) -- the tables live in HBM, the work is
// organised so that each phase runs near the unit that bounds it:
//
//   k_emit_full_rb<D>   expected log-likelihoods (pybasicbayes/distributions.py:351-366 looped
//                       over states at hmmsgd_metaobs.py:508-509) in float64 with the observation
//                       rows held in REGISTERS (two rows per thread) and the triangular factor of
//                       one state streamed through shared memory as 128-bit broadcast loads: one
//                       LDS.128 feeds four DFMA, so the kernel is bound by the FP64 pipe and not by
//                       the load/store unit; the row maximum and b = exp(ll - max) are fused in.
//   k_chain_wide        forward (hmmsgd_metaobs.py:775-803) and backward (:828-855) recursions,
//                       one WARP per (window, direction), both directions concurrently; each lane
//                       keeps two columns of the transition matrix in registers, the K-vector is
//                       exchanged through a double-buffered shared-memory slot, and instead of a
//                       per-step normaliser the messages are rescaled by exact powers of two
//                       (same deadbeat scheme as fused.cuh) with the vector maximum taken by one
//                       REDUX instruction off the dependent chain.
//   k_marginals_wide    q = norm(alpha * beta) (:516-519) + per-row log normalisers.
//   k_stats_sym         transition + NIW statistics (:873-904, util.py:73-83) as one register-
//                       blocked (8x8 per thread) contraction; only the upper triangle of
//                       sum_t q x x^T is formed (it is symmetric) and mirrored by the finaliser.
#pragma once
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// emissions
// ------------------------------------------------------------------------------------------------
#define ERB_NT 128
#define ERB_KC 8              // states staged per shared-memory refill

// padded packed-lower layout of one state's factor: row i starts at an even offset
__host__ __device__ constexpr int erb_off(int i) { return (i & 1) ? 2 * ((i >> 1) + 1) * ((i >> 1) + 1) : 2 * (i >> 1) * ((i >> 1) + 1); }
__host__ __device__ constexpr int erb_len(int D) { return erb_off(D); }   // doubles per state (D even or odd)

template <int D>
__global__ void __launch_bounds__(ERB_NT)
k_emit_full_rb(int64_t R, int T, int K, const void* __restrict__ obs, int dtype,
               const uint8_t* __restrict__ mask, const int64_t* __restrict__ starts, int mask_ll,
               const double* __restrict__ Rs, const double* __restrict__ gk, const double* __restrict__ ck,
               double* __restrict__ ll, float* __restrict__ bout, double* __restrict__ mx) {
  constexpr int PL = erb_len(D);                 // padded triangle
  constexpr int SL = PL + D + (D & 1);           // + gk (kept 16-byte aligned)
  constexpr int tri = D * (D + 1) / 2;
  extern __shared__ __align__(16) double esm[];  // [2][ERB_KC][SL]
  const int tid = threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.x * (2 * ERB_NT) + tid, r1 = r0 + ERB_NT;
  double x0[D], x1[D];
  bool dead0 = r0 >= R, dead1 = r1 >= R;
  {
    const int64_t rr[2] = {r0, r1};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      double* x = h ? x1 : x0;
      bool dead = rr[h] >= R;
      if (!dead) {
        const int b = (int)(rr[h] / T); const int t = (int)(rr[h] - (int64_t)b * T);
        const int64_t gi = starts[b] + t;
        if (mask_ll && mask && mask[gi]) dead = true;
        if (dtype == SVIHMM_F32 && (D % 4) == 0 && ((((uintptr_t)obs) & 15) == 0)) {
          const float4* p = reinterpret_cast<const float4*>((const float*)obs + gi * D);
#pragma unroll
          for (int d = 0; d < D; d += 4) { const float4 q = __ldg(p + d / 4); x[d] = q.x; x[d + 1] = q.y; x[d + 2] = q.z; x[d + 3] = q.w; }
        } else {
#pragma unroll
          for (int d = 0; d < D; ++d) x[d] = ld_obs(obs, dtype, gi * D + d);
        }
#pragma unroll
        for (int d = 0; d < D; ++d) if (isnan(x[d])) dead = true;
      } else {
#pragma unroll
        for (int d = 0; d < D; ++d) x[d] = 0.0;
      }
      if (h) dead1 = dead; else dead0 = dead;
    }
  }
  double m0 = -INFINITY, m1 = -INFINITY;
  // The factors of ERB_KC states at a time live in shared memory in the padded layout, double
  // buffered: the refill of the next tile is issued with cp.async (8 bytes per element, remapped on
  // the fly from the packed-lower layout of the global step) BEFORE the current tile is consumed, so
  // its L2 latency hides behind ~70 k FP64-pipe cycles of work (ncu on the synchronous version:
  // long-scoreboard stalls at every refill, FP64 pipe 42 % active).
  auto stage = [&](const int kc_, double* buf) {
    const int nk_ = min(ERB_KC, K - kc_);
    for (int idx = tid; idx < nk_ * D * (D + 1); idx += ERB_NT) {    // (state, i, j), j <= i
      const int s = idx / (D * (D + 1)), e = idx - s * (D * (D + 1));
      const int i = e / (D + 1), j = e - i * (D + 1);
      if (j <= i) {
        const unsigned sa = (unsigned)__cvta_generic_to_shared(buf + s * SL + erb_off(i) + j);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(Rs + (size_t)(kc_ + s) * tri + (size_t)i * (i + 1) / 2 + j) : "memory");
      }
    }
    for (int idx = tid; idx < nk_ * D; idx += ERB_NT) {
      const int s = idx / D, d = idx - s * D;
      const unsigned sa = (unsigned)__cvta_generic_to_shared(buf + s * SL + PL + d);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gk + (size_t)(kc_ + s) * D + d) : "memory");
    }
    if (nk_ & 1)                                                   // an odd tail state is paired with an all-zero factor
      for (int idx = tid; idx < SL; idx += ERB_NT) buf[nk_ * SL + idx] = 0.0;
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int idx = tid; idx < 2 * ERB_KC * SL; idx += ERB_NT) esm[idx] = 0.0;   // pads of both buffers stay zero
  __syncthreads();
  stage(0, esm);
  int cur = 0;
  for (int kc = 0; kc < K; kc += ERB_KC, cur ^= 1) {
    const int nk = min(ERB_KC, K - kc);
    const double* tile = esm + (size_t)cur * ERB_KC * SL;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                 // tile kc has landed; everyone is done with the other buffer
    if (kc + ERB_KC < K) stage(kc + ERB_KC, esm + (size_t)(cur ^ 1) * ERB_KC * SL);
#pragma unroll 1
    for (int s = 0; s < nk; s += 2) {
      const double* RA = tile + s * SL;
      const double* RB = RA + SL;
      const double2* RA2 = reinterpret_cast<const double2*>(RA);
      const double2* RB2 = reinterpret_cast<const double2*>(RB);
      double aA0 = 0.0, aA1 = 0.0, aB0 = 0.0, aB1 = 0.0;      // (state A, B) x (row 0, 1)
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const double gA = RA[PL + i], gB = RB[PL + i];
        double sA0 = -gA, sA1 = -gA, sB0 = -gB, sB1 = -gB;
#pragma unroll
        for (int j = 0; j <= i; j += 2) {
          const double2 ra = RA2[(erb_off(i) + j) / 2], rb = RB2[(erb_off(i) + j) / 2];
          sA0 = fma(ra.x, x0[j], sA0); sA1 = fma(ra.x, x1[j], sA1);
          sB0 = fma(rb.x, x0[j], sB0); sB1 = fma(rb.x, x1[j], sB1);
          if (j + 1 <= i) {
            sA0 = fma(ra.y, x0[j + 1], sA0); sA1 = fma(ra.y, x1[j + 1], sA1);
            sB0 = fma(rb.y, x0[j + 1], sB0); sB1 = fma(rb.y, x1[j + 1], sB1);
          }
        }
        aA0 = fma(sA0, sA0, aA0); aA1 = fma(sA1, sA1, aA1);
        aB0 = fma(sB0, sB0, aB0); aB1 = fma(sB1, sB1, aB1);
      }
      const int k = kc + s;
      const bool two = s + 1 < nk;
      const double cA = ck[k], cB = two ? ck[k + 1] : 0.0;
      const double la0 = dead0 ? 0.0 : cA - aA0, la1 = dead1 ? 0.0 : cA - aA1;
      const double lb0 = dead0 ? 0.0 : cB - aB0, lb1 = dead1 ? 0.0 : cB - aB1;
      m0 = fmax(m0, la0); m1 = fmax(m1, la1);
      if (two) { m0 = fmax(m0, lb0); m1 = fmax(m1, lb1); }
      if (two && !(K & 1)) {                         // K even, k even: 16-byte aligned pair
        if (r0 < R) *reinterpret_cast<double2*>(ll + r0 * K + k) = make_double2(la0, lb0);
        if (r1 < R) *reinterpret_cast<double2*>(ll + r1 * K + k) = make_double2(la1, lb1);
      } else {
        if (r0 < R) { ll[r0 * K + k] = la0; if (two) ll[r0 * K + k + 1] = lb0; }
        if (r1 < R) { ll[r1 * K + k] = la1; if (two) ll[r1 * K + k + 1] = lb1; }
      }
    }
  }
  if (!bout) return;                              // mixtures: only the component log-likelihoods are wanted
  // b = exp(ll - max): this thread re-reads the rows it has just written (L1/L2 hits)
  const int64_t rr[2] = {r0, r1};
  const double mm[2] = {m0, m1};
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (rr[h] >= R) continue;
    const double* lp = ll + rr[h] * K;
    float* bp = bout + rr[h] * K;
    if (!(K & 3) && ((((uintptr_t)bout) & 15) == 0)) {
      for (int k = 0; k < K; k += 4) {
        const double2 u = *reinterpret_cast<const double2*>(lp + k), v = *reinterpret_cast<const double2*>(lp + k + 2);
        *reinterpret_cast<float4*>(bp + k) = make_float4(__expf((float)(u.x - mm[h])), __expf((float)(u.y - mm[h])),
                                                         __expf((float)(v.x - mm[h])), __expf((float)(v.y - mm[h])));
      }
    } else {
      for (int k = 0; k < K; ++k) bp[k] = __expf((float)(lp[k] - mm[h]));
    }
    mx[rr[h]] = mm[h];
  }
}

// Diagonal emissions for many states (BASELINE config 4: K = 256, D = 64), the counterpart of
// k_emit_full_rb: thread = observation row with the row in REGISTERS (float64, converted once - the float ->
// double conversion runs at 16 lanes per clock, and k_emit_diag_tiled paid it per (row, state, d)), the
// parameters of EDR_KC states at a time in shared memory as (c2, c1) pairs, double buffered with 16-byte
// cp.async; ll = ck' + sum_d x_d (c2 x_d + c1) in the expanded form of the fused kernels (global.cuh: par2 /
// ckp), two DFMA per 128-bit broadcast load; the row maximum and b = exp(ll - max) are fused in.
#define EDR_NT 128
#define EDR_KC 16
template <int D>
__global__ void __launch_bounds__(EDR_NT)
k_emit_diag_rb(int64_t R, int T, int K, const void* __restrict__ obs, int dtype,
               const uint8_t* __restrict__ mask, const int64_t* __restrict__ starts, int mask_ll,
               const double* __restrict__ par2, const double* __restrict__ ckp,
               double* __restrict__ ll, float* __restrict__ bout, double* __restrict__ mx) {
  extern __shared__ __align__(16) double esm[];          // [2][EDR_KC][D] double2
  double2* tiles = reinterpret_cast<double2*>(esm);
  const int tid = threadIdx.x;
  const int64_t r = (int64_t)blockIdx.x * EDR_NT + tid;
  double x[D];
  bool dead = r >= R;
  if (!dead) {
    const int b = (int)(r / T); const int t = (int)(r - (int64_t)b * T);
    const int64_t gi = starts[b] + t;
    if (mask_ll && mask && mask[gi]) dead = true;
    if (dtype == SVIHMM_F32 && (D % 4) == 0 && ((((uintptr_t)obs) & 15) == 0)) {
      const float4* p = reinterpret_cast<const float4*>((const float*)obs + gi * D);
#pragma unroll
      for (int d = 0; d < D; d += 4) { const float4 q = __ldg(p + d / 4); x[d] = q.x; x[d + 1] = q.y; x[d + 2] = q.z; x[d + 3] = q.w; }
    } else {
#pragma unroll
      for (int d = 0; d < D; ++d) x[d] = ld_obs(obs, dtype, gi * D + d);
    }
#pragma unroll
    for (int d = 0; d < D; ++d) if (isnan(x[d])) dead = true;
  }
  if (dead) {
#pragma unroll
    for (int d = 0; d < D; ++d) x[d] = 0.0;
  }
  // par2 is [d][K] pairs: the tile of states [kc, kc + EDR_KC) lands as [state][d]
  auto stage = [&](const int kc, double2* buf) {
    for (int idx = tid; idx < EDR_KC * D; idx += EDR_NT) {
      const int d = idx / EDR_KC, s = idx - d * EDR_KC;
      const unsigned sa = (unsigned)__cvta_generic_to_shared(buf + s * D + d);
      if (kc + s < K)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(par2 + 2 * ((size_t)d * K + kc + s)) : "memory");
      else buf[s * D + d] = make_double2(0.0, 0.0);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  stage(0, tiles);
  double m = -INFINITY;
  int cur = 0;
  for (int kc = 0; kc < K; kc += EDR_KC, cur ^= 1) {
    const double2* tile = tiles + (size_t)cur * EDR_KC * D;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                 // tile kc has landed; everyone is done with the other buffer
    if (kc + EDR_KC < K) stage(kc + EDR_KC, tiles + (size_t)(cur ^ 1) * EDR_KC * D);
#pragma unroll 1
    for (int s = 0; s < EDR_KC; s += 2) {
      const double2* pa = tile + s * D; const double2* pb = pa + D;
      double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;           // two states x two partial sums
#pragma unroll
      for (int d = 0; d < D; d += 2) {
        const double2 u0 = pa[d], u1 = pa[d + 1], v0 = pb[d], v1 = pb[d + 1];
        a0 = fma(x[d], fma(u0.x, x[d], u0.y), a0); a1 = fma(x[d + 1], fma(u1.x, x[d + 1], u1.y), a1);
        b0 = fma(x[d], fma(v0.x, x[d], v0.y), b0); b1 = fma(x[d + 1], fma(v1.x, x[d + 1], v1.y), b1);
      }
      const int k = kc + s;
      if (k < K) {
        const double la = dead ? 0.0 : ckp[k] + (a0 + a1);
        m = fmax(m, la);
        if (r < R) ll[r * K + k] = la;
      }
      if (k + 1 < K) {
        const double lb = dead ? 0.0 : ckp[k + 1] + (b0 + b1);
        m = fmax(m, lb);
        if (r < R) ll[r * K + k + 1] = lb;
      }
    }
  }
  if (!bout || r >= R) return;                  // mixtures: only the component log-likelihoods are wanted
  // b = exp(ll - max): this thread re-reads the row it has just written (L1/L2 hits)
  const double* lp = ll + r * K;
  float* bp = bout + r * K;
  if (!(K & 3) && ((((uintptr_t)bout) & 15) == 0)) {
    for (int k = 0; k < K; k += 4) {
      const double2 u = *reinterpret_cast<const double2*>(lp + k), v = *reinterpret_cast<const double2*>(lp + k + 2);
      *reinterpret_cast<float4*>(bp + k) = make_float4(__expf((float)(u.x - m)), __expf((float)(u.y - m)),
                                                       __expf((float)(v.x - m)), __expf((float)(v.y - m)));
    }
  } else {
    for (int k = 0; k < K; ++k) bp[k] = __expf((float)(lp[k] - m));
  }
  mx[r] = m;
}

// ------------------------------------------------------------------------------------------------
// recursions: one warp per (window, direction)
// ------------------------------------------------------------------------------------------------
#define CW_XTB 157            // target biased exponent of the vector maximum: 2^30 (as fused.cuh)
#define CW_PF 4               // steps of b prefetched in registers
#define CW_L2 12              // further steps prefetched into L2

__device__ __forceinline__ void cw_ffma2(unsigned long long& acc, const unsigned long long a, const unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long cw_pack(const float x, const float y) {
  return (unsigned long long)__float_as_uint(x) | ((unsigned long long)__float_as_uint(y) << 32);
}
__device__ __forceinline__ float cw_hsum(const unsigned long long v) {
  return __uint_as_float((unsigned)v) + __uint_as_float((unsigned)(v >> 32));
}

// One step of a chain.  v0/v1: this lane's two components of the carried vector (alpha~ forward,
// b*beta~ backward).  With e(v) the biased exponent of max_j v[j], the exponent shift of step s is
//     d_s = e(v_{s-2}) - XTB - d_{s-1}        (fused.cuh: the shift already in flight is subtracted)
// so the measurement (one REDUX per step) is never on the dependent chain.
template <bool FWD>
__device__ __forceinline__ void cw_step(float& v0, float& v1, const unsigned long long (&c0)[32],
                                        const unsigned long long (&c1)[32], const float b0, const float b1,
                                        const float* slot_r, float* slot_w, const int lane,
                                        float* __restrict__ ow, const int off, const bool ac0, const bool ac1,
                                        int* __restrict__ ew, const int t,
                                        int& e1, int& e2, int& da, int& E) {
  __syncwarp();                                        // the vector of the previous step is in slot_r
  int d = e2 - CW_XTB - da;
  d = max(-60, min(60, d));
  const float r = __uint_as_float((unsigned)(127 - d) << 23);
  unsigned long long a0 = 0ull, a1 = 0ull, g0 = 0ull, g1 = 0ull;
  const ulonglong2* x2 = reinterpret_cast<const ulonglong2*>(slot_r);
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const ulonglong2 x = x2[q];                        // components 4q..4q+3 as two packed pairs
    cw_ffma2(a0, x.x, c0[2 * q]); cw_ffma2(g0, x.x, c1[2 * q]);
    cw_ffma2(a1, x.y, c0[2 * q + 1]); cw_ffma2(g1, x.y, c1[2 * q + 1]);
  }
  const float mA = (cw_hsum(a0) + cw_hsum(a1)) * r, mG = (cw_hsum(g0) + cw_hsum(g1)) * r;
  v0 = mA * b0; v1 = mG * b1;
  slot_w[lane] = v0; slot_w[lane + 32] = v1;           // the only stores on the dependent chain
  E += d;
  if (ac0) ow[off] = FWD ? v0 : mA;
  if (ac1) ow[off + 32] = FWD ? v1 : mG;
  if (FWD && lane == 0) ew[t] = E;
  const unsigned mu = __reduce_max_sync(0xffffffffu, max(__float_as_uint(v0), __float_as_uint(v1)));
  e2 = e1; e1 = (int)(mu >> 23); da = d;
}

// The chain of one (window, direction): M[i][j] with out[j] = sum_i in[i] M[i][j] (forward M = P,
// backward M = P^T).  bw/ow: b and output rows of this window; every per-step address is the
// window base plus ONE 32-bit per-lane offset (component 1 sits 32 floats further).
template <bool FWD>
__device__ __forceinline__ void cw_run(const int T, const int K, const float* __restrict__ M,
                                       const float* __restrict__ pi0, const float* __restrict__ bw,
                                       float* __restrict__ ow, int* __restrict__ ew, float* sA, float* sB,
                                       const int lane) {
  const int j1 = lane + 32;
  const bool ac0 = lane < K, ac1 = j1 < K;
  unsigned long long c0[32], c1[32];                   // (M[i][j], M[i+1][j]) pairs of this lane's two columns
#pragma unroll
  for (int i = 0; i < 64; i += 2) {
    const float p00 = (i < K && ac0) ? __ldg(M + i * K + lane) : 0.f;
    const float p01 = (i + 1 < K && ac0) ? __ldg(M + (i + 1) * K + lane) : 0.f;
    const float p10 = (i < K && ac1) ? __ldg(M + i * K + j1) : 0.f;
    const float p11 = (i + 1 < K && ac1) ? __ldg(M + (i + 1) * K + j1) : 0.f;
    c0[i / 2] = cw_pack(p00, p01); c1[i / 2] = cw_pack(p10, p11);
  }
  const int dk = FWD ? K : -K;
  int t = FWD ? 0 : T - 1;
  int off = t * K + lane;                              // row t, component 0
  float v0, v1;
  {
    const float b0 = ac0 ? bw[off] : 0.f, b1 = ac1 ? bw[off + 32] : 0.f;
    if (FWD) { v0 = ac0 ? __ldg(pi0 + lane) * b0 : 0.f; v1 = ac1 ? __ldg(pi0 + j1) * b1 : 0.f; }
    else { v0 = b0; v1 = b1; }
    if (ac0) ow[off] = FWD ? v0 : 1.f;
    if (ac1) ow[off + 32] = FWD ? v1 : 1.f;
    if (FWD && lane == 0) ew[t] = 0;
  }
  sA[lane] = v0; sA[j1] = v1;                          // step 1 reads slot A
  int e1, e2, da = 0, E = 0;
  {
    const unsigned mu = __reduce_max_sync(0xffffffffu, max(__float_as_uint(v0), __float_as_uint(v1)));
    e1 = (int)(mu >> 23); e2 = e1;                     // first shift: measurement of v_0, nothing in flight
  }
  float pb0[CW_PF], pb1[CW_PF];
  int pf = off;                                        // offset of the row being prefetched
#pragma unroll
  for (int u = 0; u < CW_PF; ++u) {
    pf += dk;
    const bool ok = 1 + u < T;
    pb0[u] = (ok && ac0) ? bw[pf] : 0.f;
    pb1[u] = (ok && ac1) ? bw[pf + 32] : 0.f;
  }
  const int dt = FWD ? 1 : -1;
  int s = 1;
  for (; s + CW_PF <= T; s += CW_PF) {                 // full groups: steps s .. s+PF-1
#pragma unroll
    for (int u = 0; u < CW_PF; ++u) {
      const float b0 = pb0[u], b1 = pb1[u];
      pf += dk;
      if (s + u + CW_PF + CW_L2 < T) {                 // pull the row of step s+u+PF+L2 towards L2/L1 (HBM latency
        if (ac0) asm volatile("prefetch.global.L2 [%0];" ::"l"(bw + pf + CW_L2 * dk)); // is ~5 steps long)
        if (ac1) asm volatile("prefetch.global.L2 [%0];" ::"l"(bw + pf + CW_L2 * dk + 32));
      }
      const bool more = s + u + CW_PF < T;             // refill this prefetch slot for step s+u+PF
      pb0[u] = (more && ac0) ? bw[pf] : 0.f;
      pb1[u] = (more && ac1) ? bw[pf + 32] : 0.f;
      off += dk; t += dt;
      if (u & 1) cw_step<FWD>(v0, v1, c0, c1, b0, b1, sB, sA, lane, ow, off, ac0, ac1, ew, t, e1, e2, da, E);
      else       cw_step<FWD>(v0, v1, c0, c1, b0, b1, sA, sB, lane, ow, off, ac0, ac1, ew, t, e1, e2, da, E);
    }
  }
  // tail: fewer than PF steps left; their b values are already in the prefetch registers.  The group
  // loop always ends on an even number of steps, so the parity of the slots restarts at A.
#pragma unroll
  for (int u = 0; u < CW_PF - 1; ++u) {
    if (s + u < T) {
      off += dk; t += dt;
      if (u & 1) cw_step<FWD>(v0, v1, c0, c1, pb0[u], pb1[u], sB, sA, lane, ow, off, ac0, ac1, ew, t, e1, e2, da, E);
      else       cw_step<FWD>(v0, v1, c0, c1, pb0[u], pb1[u], sA, sB, lane, ow, off, ac0, ac1, ew, t, e1, e2, da, E);
    }
  }
}

// alpha_out[r][k]: forward messages scaled by 2^-E[r]; beta_out[r][k]: backward messages (arbitrary
// power-of-two scale per row: the marginals are normalised per row).  One warp per (window,
// direction); a CTA holds the four chains of two windows.
__global__ void __launch_bounds__(128)
k_chain_wide(int B, int T, int K, const float* __restrict__ P, const float* __restrict__ PT,
             const float* __restrict__ pi0, const float* __restrict__ b,
             float* __restrict__ alpha_out, float* __restrict__ beta_out, int* __restrict__ E_out) {
  __shared__ __align__(16) float slot[4][2][64];
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int chain = blockIdx.x * 4 + wp;
  const int w = chain >> 1;
  if (w >= B) return;
  const size_t base = (size_t)w * T * K;
  if (!(chain & 1)) cw_run<true>(T, K, P, pi0, b + base, alpha_out + base, E_out + (size_t)w * T, slot[wp][0], slot[wp][1], lane);
  else cw_run<false>(T, K, PT, pi0, b + base, beta_out + base, nullptr, slot[wp][0], slot[wp][1], lane);
}

// q[r] = alpha[r]*beta[r] / sum;  lt[r] = log(sum_k alpha[r][k]) + E[r] ln 2  ( = logsumexp_k lalpha[t,k]
// minus the running sum of row maxima).  One warp per row, grid-stride.
__global__ void __launch_bounds__(256)
k_marginals_wide(int64_t R, int K, const float* __restrict__ alpha, const float* __restrict__ beta,
                 const int* __restrict__ E, float* __restrict__ q, double* __restrict__ lt) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int k0 = lane, k1 = lane + 32;
  for (int64_t r = w0; r < R; r += nw) {
    const float al0 = k0 < K ? alpha[r * K + k0] : 0.f, al1 = k1 < K ? alpha[r * K + k1] : 0.f;
    const float p0 = k0 < K ? al0 * beta[r * K + k0] : 0.f, p1 = k1 < K ? al1 * beta[r * K + k1] : 0.f;
    float sa = al0 + al1, sp = p0 + p1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sa += __shfl_xor_sync(0xffffffffu, sa, o);
      sp += __shfl_xor_sync(0xffffffffu, sp, o);
    }
    const float inv = 1.f / sp;
    if (k0 < K) q[r * K + k0] = p0 * inv;
    if (k1 < K) q[r * K + k1] = p1 * inv;
    if (lane == 0) lt[r] = (double)logf(sa) + (double)E[r] * M_LN2;
  }
}

// per window: seq[2b] = logZ = lt[T-1] + sum_t mx[t];  seq[2b+1] = Q4 bound (hmmsgd_metaobs.py:257-271)
//   = sum_t (lt[t] + sum_{s<=t} mx[s]) = sum_t lt[t] + sum_t (T - t) mx[t]
__global__ void __launch_bounds__(256)
k_seq_logz_lt(int B, int T, const double* __restrict__ lt, const double* __restrict__ mx, double* __restrict__ seq) {
  const int lane = threadIdx.x & 31;
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (s >= B) return;
  double smx = 0.0, q4 = 0.0;
  for (int t = lane; t < T; t += 32) {
    const double m = mx[(size_t)s * T + t];
    smx += m;
    q4 += lt[(size_t)s * T + t] + (double)(T - t) * m;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    smx += __shfl_xor_sync(0xffffffffu, smx, o);
    q4 += __shfl_xor_sync(0xffffffffu, q4, o);
  }
  if (lane == 0) { seq[2 * s] = lt[(size_t)s * T + T - 1] + smx; seq[2 * s + 1] = q4; }
}

// ------------------------------------------------------------------------------------------------
// statistics
// ------------------------------------------------------------------------------------------------
#define SS_NT 128
#define SS_RC 32              // rows per shared-memory stage
#define SS_TN 128             // feature columns per CTA
#define SS_XS 68              // row stride of the staged observations (D <= 64, +1 for the constant 1)

struct StatsSymArgs {
  int B, T, K, D, NF, diag, wrap, dtype;
  int64_t R, rows_per_split;
  const float* q; const void* obs; const uint8_t* mask; const int64_t* starts;
  float* part;               // [nsplit][K][NF]
};

// columns: [ next q (K) | w | w x_d (D) | w x_i x_j, i <= j row-major (D(D+1)/2)  or  w x_d^2 (D, diag) ]
__global__ void __launch_bounds__(SS_NT) k_stats_sym(const StatsSymArgs a) {
  __shared__ __align__(16) float Ls[SS_RC][64];
  __shared__ __align__(16) float Fs[SS_RC][SS_TN];
  __shared__ float xs[SS_RC][SS_XS];
  __shared__ float wv[SS_RC];
  const int tid = threadIdx.x, tn = tid & 15, tm = tid >> 4;
  const int K = a.K, D = a.D, T = a.T;
  const int n0 = blockIdx.x * SS_TN;
  const int64_t rbeg = (int64_t)blockIdx.y * a.rows_per_split;
  const int64_t rend = min(a.R, rbeg + a.rows_per_split);
  // descriptor of the feature column this thread generates: val = w * xs[d1] * xs[d2], xs[D] = 1;
  // kind 0 = next q
  const int col = n0 + tid;
  int kind = -1, d1 = D, d2 = D;
  if (col < K) kind = 0;
  else if (col < a.NF) {
    kind = 1;
    int c = col - K;
    if (c == 0) { d1 = D; d2 = D; }
    else if (c <= D) { d1 = c - 1; d2 = D; }
    else {
      c -= D + 1;
      if (a.diag) { d1 = c; d2 = c; }
      else { int i = 0; while (c >= D - i) { c -= D - i; ++i; } d1 = i; d2 = i + c; }
    }
  }
  const bool need_x = n0 + SS_TN > K;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  __shared__ int64_t grow[SS_RC], nrow[SS_RC];   // series row of (b,t) / table row of (b,t+1), -1 = none
  for (int64_t rc = rbeg; rc < rend; rc += SS_RC) {
    if (tid < SS_RC) {
      const int64_t r = rc + tid;
      int64_t g = -1, nx = -1;
      if (r < rend) {
        const int b = (int)(r / T); const int t = (int)(r - (int64_t)b * T);
        g = a.starts[b] + t;
        if (t + 1 < T) nx = r + 1; else if (a.wrap) nx = (int64_t)b * T;
      }
      grow[tid] = g; nrow[tid] = nx;
    }
    for (int idx = tid; idx < SS_RC * 64; idx += SS_NT) {
      const int rr = idx >> 6, m = idx & 63;
      const int64_t r = rc + rr;
      Ls[rr][m] = (r < rend && m < K) ? a.q[r * K + m] : 0.f;
    }
    __syncthreads();
    if (need_x) {
      for (int idx = tid; idx < SS_RC * D; idx += SS_NT) {
        const int rr = idx / D, d = idx - rr * D;
        const int64_t g = grow[rr];
        xs[rr][d] = g >= 0 ? (float)ld_obs(a.obs, a.dtype, g * D + d) : 0.f;
      }
      __syncthreads();
      if (tid < SS_RC) {
        const int64_t g = grow[tid];
        float w = 0.f;
        if (g >= 0) {
          w = (a.mask && a.mask[g]) ? 0.f : 1.f;
          bool bad = false;
          for (int d = 0; d < D; ++d) bad |= isnan(xs[tid][d]);
          if (bad || w == 0.f) { w = 0.f; for (int d = 0; d < D; ++d) xs[tid][d] = 0.f; }
        }
        wv[tid] = w;
        xs[tid][D] = 1.f;
      }
      __syncthreads();
    }
    if (kind == 0) {
#pragma unroll 8
      for (int rr = 0; rr < SS_RC; ++rr) {
        const int64_t nx = nrow[rr];
        Fs[rr][tid] = nx >= 0 ? a.q[nx * K + col] : 0.f;
      }
    } else if (kind == 1) {
#pragma unroll 8
      for (int rr = 0; rr < SS_RC; ++rr) Fs[rr][tid] = wv[rr] * xs[rr][d1] * xs[rr][d2];
    } else {
      for (int rr = 0; rr < SS_RC; ++rr) Fs[rr][tid] = 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int rr = 0; rr < SS_RC; ++rr) {
      const float4 l0 = *reinterpret_cast<const float4*>(&Ls[rr][tm * 4]);
      const float4 l1 = *reinterpret_cast<const float4*>(&Ls[rr][32 + tm * 4]);
      const float4 f0 = *reinterpret_cast<const float4*>(&Fs[rr][tn * 4]);
      const float4 f1 = *reinterpret_cast<const float4*>(&Fs[rr][64 + tn * 4]);
      const float l[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
      const float f[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(l[i], f[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = (i < 4 ? tm * 4 + i : 32 + tm * 4 + (i - 4));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tn * 4 + j : 64 + tn * 4 + (j - 4));
      if (m < K && n < a.NF) a.part[((size_t)blockIdx.y * K + m) * a.NF + n] = acc[i][j];
    }
  }
}

// Tensor-core version of k_stats_sym: the same contraction Out[m][n] = sum_r q[r][m] F[r][n] with
// mma.m16n8k8 TF32, every operand split hi/lo (3xTF32, float32 accumulators), so the statistics keep
// float32-level accuracy (the 1e-5 parity bound) at ~1/3 of the issue slots of the FFMA version.
// CTA = 4 warps, warp w owns the 16-state tile w; the CTA covers SM_NT feature tiles of 8 columns and
// one row split.  Per stage of 32 rows the feature columns are generated ONCE per CTA (already split
// into TF32 hi/lo words) into shared memory with a row stride of 136 words, and q with a stride of 72,
// which makes every fragment load conflict-free (bank = 8*tig + g).
#define SM_NT 16              // feature tiles (of 8 columns) per CTA
#define SM_FS 136             // row stride of the feature stage (words)
#define SM_QS 72              // row stride of the q stage (words)
#define SM_SMEM (4 * SS_RC * (2 * SM_QS + 2 * SM_FS + 2 * SS_XS) + 2 * 16 * SS_RC + 4 * SS_RC)
__device__ __forceinline__ void sm_split(const float v, unsigned& hi, unsigned& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
  const float r = v - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void sm_mma(float (&d)[4], const unsigned (&a)[4], const unsigned b0, const unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(SS_NT) k_stats_mma(const StatsSymArgs a) {
  // dynamic shared memory (SM_SMEM bytes): q rows and observation rows of TWO stages (the next stage is
  // fetched with cp.async while the current one is consumed), the split feature stage, per-row scalars
  extern __shared__ __align__(16) unsigned char smm[];
  float (*Qs)[SS_RC][SM_QS] = reinterpret_cast<float (*)[SS_RC][SM_QS]>(smm);
  float (*xs)[SS_RC][SS_XS] = reinterpret_cast<float (*)[SS_RC][SS_XS]>(Qs + 2);
  unsigned (*Fh)[SM_FS] = reinterpret_cast<unsigned (*)[SM_FS]>(xs + 2);
  unsigned (*Fl)[SM_FS] = Fh + SS_RC;
  int64_t (*grow)[SS_RC] = reinterpret_cast<int64_t (*)[SS_RC]>(Fl + SS_RC);
  int64_t (*nrow)[SS_RC] = grow + 2;
  float* wv = reinterpret_cast<float*>(nrow + 2);
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5, g = lane >> 2, tig = lane & 3;
  const int K = a.K, D = a.D, T = a.T;
  const int n0 = blockIdx.x * (SM_NT * 8);
  const int64_t rbeg = (int64_t)blockIdx.y * a.rows_per_split;
  const int64_t rend = min(a.R, rbeg + a.rows_per_split);
  const int col = n0 + tid;                       // the feature column this thread generates (128 per CTA)
  int kind = -1, d1 = D, d2 = D;
  if (col < K) kind = 0;
  else if (col < a.NF) {
    kind = 1;
    int c = col - K;
    if (c == 0) { d1 = D; d2 = D; }
    else if (c <= D) { d1 = c - 1; d2 = D; }
    else {
      c -= D + 1;
      if (a.diag) { d1 = c; d2 = c; }
      else { int i = 0; while (c >= D - i) { c -= D - i; ++i; } d1 = i; d2 = i + c; }
    }
  }
  const bool need_x = n0 + SM_NT * 8 > K;
  const bool mact = 16 * wp < K;                  // this warp's state tile exists
  const bool xvec = a.dtype == SVIHMM_F32 && (D & 3) == 0 && ((((uintptr_t)a.obs) & 15) == 0);
  const bool qvec = (K & 3) == 0 && ((((uintptr_t)a.q) & 15) == 0);
  float acc[SM_NT][4];
#pragma unroll
  for (int j = 0; j < SM_NT; ++j)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[j][c] = 0.f;

  // fetch stage `rc` into buffer `bf`: row indices with plain stores, q / x rows with cp.async
  auto fetch = [&](const int64_t rc, const int bf) {
    for (int idx = tid; idx < SS_RC * 16; idx += SS_NT) {          // q rows: 16 float4 per row (zero padded)
      const int rr = idx >> 4, m4 = (idx & 15) * 4;
      const int64_t r = rc + rr;
      float* dst = &Qs[bf][rr][m4];
      if (r < rend && qvec && m4 + 3 < K) {
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(a.q + r * K + m4) : "memory");
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) dst[u] = (r < rend && m4 + u < K) ? a.q[r * K + m4 + u] : 0.f;
      }
    }
    if (tid < SS_RC) {
      const int64_t r = rc + tid;
      int64_t gg = -1, nx = -1;
      if (r < rend) {
        const int b = (int)(r / T); const int t = (int)(r - (int64_t)b * T);
        gg = a.starts[b] + t;
        if (t + 1 < T) nx = r + 1; else if (a.wrap) nx = (int64_t)b * T;
      }
      grow[bf][tid] = gg; nrow[bf][tid] = nx;
    }
    if (need_x) {
      // every thread recomputes the series row of the rows it fetches (grow of this stage is not visible yet)
      if (xvec) {
        const int v4 = D >> 2;
        for (int idx = tid; idx < SS_RC * v4; idx += SS_NT) {
          const int rr = idx / v4, d4 = (idx - rr * v4) * 4;
          const int64_t r = rc + rr;
          float* dst = &xs[bf][rr][d4];
          if (r < rend) {
            const int b = (int)(r / T); const int t = (int)(r - (int64_t)b * T);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)),
                         "l"((const float*)a.obs + (a.starts[b] + t) * D + d4) : "memory");
          } else { dst[0] = dst[1] = dst[2] = dst[3] = 0.f; }
        }
      } else {
        for (int idx = tid; idx < SS_RC * D; idx += SS_NT) {
          const int rr = idx / D, d = idx - rr * D;
          const int64_t r = rc + rr;
          float v = 0.f;
          if (r < rend) {
            const int b = (int)(r / T); const int t = (int)(r - (int64_t)b * T);
            v = (float)ld_obs(a.obs, a.dtype, (a.starts[b] + t) * D + d);
          }
          xs[bf][rr][d] = v;
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  fetch(rbeg, 0);
  int bf = 0;
  for (int64_t rc = rbeg; rc < rend; rc += SS_RC, bf ^= 1) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();                              // stage rc is in buffer bf; everyone is done with the other buffer and F
    if (rc + SS_RC < rend) fetch(rc + SS_RC, bf ^ 1);
    if (need_x) {
      if (tid < SS_RC) {
        const int64_t gg = grow[bf][tid];
        float w = 0.f;
        if (gg >= 0) {
          w = (a.mask && a.mask[gg]) ? 0.f : 1.f;
          bool bad = false;
          for (int d = 0; d < D; ++d) bad |= isnan(xs[bf][tid][d]);
          if (bad || w == 0.f) { w = 0.f; for (int d = 0; d < D; ++d) xs[bf][tid][d] = 0.f; }
        }
        wv[tid] = w;
        xs[bf][tid][D] = 1.f;
      }
      __syncthreads();
    }
    if (kind == 0) {
#pragma unroll 8
      for (int rr = 0; rr < SS_RC; ++rr) {
        const int64_t nx = nrow[bf][rr];
        unsigned hi, lo;
        sm_split(nx >= 0 ? a.q[nx * K + col] : 0.f, hi, lo);
        Fh[rr][tid] = hi; Fl[rr][tid] = lo;
      }
    } else {
      const float live = kind == 1 ? 1.f : 0.f;       // columns beyond NF: zeros
#pragma unroll 8
      for (int rr = 0; rr < SS_RC; ++rr) {
        unsigned hi, lo;
        sm_split(live * wv[rr] * xs[bf][rr][d1] * xs[bf][rr][d2], hi, lo);
        Fh[rr][tid] = hi; Fl[rr][tid] = lo;
      }
    }
    __syncthreads();
    if (mact) {
#pragma unroll
      for (int ks = 0; ks < SS_RC / 8; ++ks) {
        const int r0 = ks * 8 + tig, r1 = r0 + 4, c0 = 16 * wp + g;
        unsigned ah[4], al[4];
        sm_split(Qs[bf][r0][c0], ah[0], al[0]); sm_split(Qs[bf][r0][c0 + 8], ah[1], al[1]);
        sm_split(Qs[bf][r1][c0], ah[2], al[2]); sm_split(Qs[bf][r1][c0 + 8], ah[3], al[3]);
        // all SM_NT tiles unconditionally (columns beyond NF hold zeros): no branches around the MMAs, and
        // the three MMAs of a tile are spread over three sweeps so that consecutive MMAs are independent
        const unsigned* fh0 = &Fh[r0][g]; const unsigned* fh1 = &Fh[r1][g];
        const unsigned* fl0 = &Fl[r0][g]; const unsigned* fl1 = &Fl[r1][g];
#pragma unroll
        for (int j = 0; j < SM_NT; ++j) sm_mma(acc[j], al, fh0[8 * j], fh1[8 * j]);
#pragma unroll
        for (int j = 0; j < SM_NT; ++j) sm_mma(acc[j], ah, fl0[8 * j], fl1[8 * j]);
#pragma unroll
        for (int j = 0; j < SM_NT; ++j) sm_mma(acc[j], ah, fh0[8 * j], fh1[8 * j]);
      }
    }
  }
  if (mact) {
#pragma unroll
    for (int j = 0; j < SM_NT; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int m = 16 * wp + g + 8 * (c >> 1), n = n0 + 8 * j + 2 * tig + (c & 1);
        if (m < K && n < a.NF) a.part[((size_t)blockIdx.y * K + m) * a.NF + n] = acc[j][c];
      }
  }
}

// Sum the row-split partials in float64, mirror the symmetric second moments, lay the statistics
// out as include/svihmm.h documents (same tail as k_stats_finalize).
__global__ void __launch_bounds__(256)
k_stats_sym_finalize(int B, int T, int K, int D, int DD, int NF, int diag, int nsplit,
                     const float* __restrict__ part, const float* __restrict__ q,
                     const double* __restrict__ seq, const double* __restrict__ prior_tran, int add_prior,
                     double* __restrict__ out, size_t slen) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= slen) return;
  const size_t o_n = (size_t)K * K, o_sx = o_n + K, o_sxx = o_sx + (size_t)K * D,
               o_q0 = o_sxx + (size_t)K * DD, o_tail = o_q0 + K;
  int m = -1, n = 0;
  double v = 0.0;
  if (idx < o_n) { m = (int)(idx / K); n = (int)(idx % K); }
  else if (idx < o_sx) { m = (int)(idx - o_n); n = K; }
  else if (idx < o_sxx) { const size_t e = idx - o_sx; m = (int)(e / D); n = K + 1 + (int)(e % D); }
  else if (idx < o_q0) {
    const size_t e = idx - o_sxx; m = (int)(e / DD);
    const int c = (int)(e % DD);
    if (diag) n = K + 1 + D + c;
    else {
      int i = c / D, j = c - i * D;
      if (i > j) { const int tmp = i; i = j; j = tmp; }
      n = K + 1 + D + i * D - i * (i - 1) / 2 + (j - i);       // row-major index of (i <= j)
    }
  }
  if (m >= 0) {
    for (int z = 0; z < nsplit; ++z) v += (double)part[((size_t)z * K + m) * NF + n];
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
