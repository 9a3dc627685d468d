// K5: the whole global step in ONE launch on the device-resident float64 master parameters:
//   block 0        transitions: natural-gradient blend (hmmsgd_metaobs.py:1029-1045) or batch
//                  update (hmmbatchcd.py:179-184), then the per-step constants: digamma transforms
//                  (hmmsgd_metaobs.py:502-504) and the stationary "initial distribution" (quirk Q3,
//                  :413-418)
//   blocks 1..     emissions: NIW natural-parameter blend (hmmsgd_metaobs.py:1048-1069 with
//                  util.py:28-60) or conjugate update (pybasicbayes/distributions.py:240-276), then
//                  the constants of the expected log-likelihood (distributions.py:351-366)
// The reference recomputes np.linalg.eig(A_mean.T) for every meta-observation; here the Perron
// vector is obtained once per global step with the Grassmann-Taksar-Heyman elimination (no
// subtractions, componentwise accurate), which equals the eigenvector eig returns up to scale.
#pragma once
#include "common.cuh"

enum { GM_PREP = 0, GM_SVI = 1, GM_BATCH = 2, GM_BSGD = 3 };   // BSGD: hmmbatchsgd.py:202-259 (SVI blend + explicit var_init)

struct GlobalArgs {
  int K, D, DD, diag, cat, mode, user_init;
  int KE, C;                                // emission components K*C, components per state
  double* ada_G;                            // non-NULL: AdaGrad-like transition step (hmmsgd_metaobs.py:1036-1040)
  double *omega, *omega_prior, *lw;         // mixtures only
  // one-shot all-reduce over NVLink peer memory (world > 1).  Exchange area of every rank:
  // [ flags: world x nb u64 | receive slots: 2 parities x world x slen doubles ]; xbase[p] = rank p's
  // area as addressed from this rank; stats = this rank's own statistics (plain device memory);
  // red_out = local buffer receiving the sums
  int world, rank, nb;
  unsigned long long* xbase[8];
  unsigned long long seq;
  double* red_out;
  const double* stats_local;                // this rank's own statistics (stats = red_out when world > 1)
  size_t slen;
  size_t plen;
  double *W, *vinit, *emit;                 // vinit: [0,K) the vector in use, [K,2K) the user-given one
  const double *prior_tran, *prior_init, *prior_emit, *stats;
  double lrate, bA, bE;
  double *gth, *rowsum, *ckc;               // scratch: 2*K*K, K, 2*K*D doubles
  float *Pt, *PtT, *pi0;
  double *Rs, *gk, *ck;
  long long* dbg;                           // optional clock64 stamps (debug)
  double *par2, *ckp;                       // diagonal, fused-kernel form: [d][k] (-Rs, 2 Rs mu) and ck - sum Rs mu^2
  int* status;                              // bit 0: a scale matrix lost positive definiteness (see svihmm_check)
  double* zero_buf;                         // non-NULL: slen doubles zeroed by block 0 (the NEXT E-step's accumulator, svihmm_svi_run)
  int gth_ext;                              // the stationary vector is left to k_gth_cluster (gth_cluster.cuh), which follows
};

// The global-step kernel runs each code path once per launch, so its time is dominated by cold
// instruction fetches (ncu: stall_no_inst); the double-precision helpers are therefore kept out of
// line (one copy each) and un-unrolled.
__device__ __noinline__ double dlog_ni(double x) { return log(x); }
__device__ __noinline__ double dexp_ni(double x) { return exp(x); }

// 1/s to full double precision without the ~120-cycle division routine: rcp.approx + Newton steps
__device__ __forceinline__ double gth_rcp(const double s) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(s));
#pragma unroll
  for (int it = 0; it < 3; ++it) { const double e = fma(-s, r, 1.0); r = fma(r, e, r); }
  return r;
}

// psi(x), float64: recurrence up to x >= 10, then the asymptotic series through B14 (truncation
// error < 5e-17 there).  x > 0 only.
__device__ __noinline__ double digamma_fast(double x) {
  double r = 0.0;
  if (!(x > 0.0)) return nan("");                   // Dirichlet / NIW parameters are positive
#pragma unroll 1
  while (x < 10.0) { r -= gth_rcp(x); x += 1.0; }       // up to 10 reciprocals on the serial tail of the kernel
  const double xi = gth_rcp(x), x2 = xi * xi;
  r += dlog_ni(x) - 0.5 * xi
     - x2 * (1.0 / 12 - x2 * (1.0 / 120 - x2 * (1.0 / 252 - x2 * (1.0 / 240
     - x2 * (1.0 / 132 - x2 * (691.0 / 32760 - x2 * (1.0 / 12)))))));
  return r;
}

struct GStats { const double *A, *n, *sx, *sxx, *q0; };
__device__ inline GStats gstats(const double* s, int K, int KE, int D, int DD) {
  GStats v;
  v.A = s; v.n = s + (size_t)K * K; v.sx = v.n + KE; v.sxx = v.sx + (size_t)KE * D; v.q0 = v.sxx + (size_t)KE * DD;
  return v;
}

// ---- fused all-reduce of the statistics (multi-GPU) ----------------------------------------------
// The windows of a minibatch are sharded over the ranks (hmmsgd_metaobs.py:405-433 only ADDS their
// results), so the global step needs sum_p stats_p.  Instead of an NCCL all-reduce followed by the
// update kernel, every block of k_global_step exchanges exactly the statistics IT consumes with the
// same block of every peer over NVLink: it PUSHES its range of the local statistics into a receive
// slot in each peer's memory (posted P2P stores: one-way latency, no round trip), fences, raises a
// per-(source, block) sequence flag in the peer's memory, waits on its OWN flags (local polling) and
// sums the received copies in rank order -- so every rank gets bitwise identical sums and the
// replicas stay in lock-step -- into red_out, which the update code then reads as usual.  Receive
// slots alternate by step parity: a peer cannot push step s+2 before this rank has raised its flags
// of step s+1, i.e. finished reading the slots of step s.
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_peer(double* p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
#define GTH_SMEM_KMAX 64
struct CommRange { size_t lo, hi; };
__device__ void comm_exchange(const GlobalArgs& a, const CommRange* rg, const int nrg) {
  const int tid = threadIdx.x, nth = blockDim.x, W = a.world, me = a.rank;
  const size_t nflag = (size_t)W * a.nb;
  const size_t par = (size_t)(a.seq & 1) * W * a.slen;
  // push this block's ranges into slot `me` of every peer
  for (int r = 0; r < nrg; ++r)
    for (size_t i = rg[r].lo + tid; i < rg[r].hi; i += nth) {
      const double v = a.stats_local[i];
      for (int p = 0; p < W; ++p)
        if (p != me) st_peer(reinterpret_cast<double*>(a.xbase[p] + nflag) + par + (size_t)me * a.slen + i, v);
    }
  __syncthreads();                       // every thread's pushes happen-before the flags below: the release at
  if (tid < W && tid != me) {            // system scope is cumulative over what the barrier made visible to its thread
    st_release_sys(a.xbase[tid] + (size_t)me * a.nb + blockIdx.x, a.seq);          // tell peer `tid`
    const unsigned long long* mine = a.xbase[me] + (size_t)tid * a.nb + blockIdx.x; // hear from peer `tid`
    while (ld_acquire_sys(mine) < a.seq) {}
  }
  __syncthreads();
  const double* rcv = reinterpret_cast<const double*>(a.xbase[me] + nflag) + par;
  for (int r = 0; r < nrg; ++r)
    for (size_t i = rg[r].lo + tid; i < rg[r].hi; i += nth) {
      double s = 0.0;
      for (int p = 0; p < W; ++p) s += p == me ? a.stats_local[i] : __ldcv(rcv + (size_t)p * a.slen + i);
      a.red_out[i] = s;
    }
}

// ---- block 0 -------------------------------------------------------------------------------
__device__ __forceinline__ void bar_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Perron vector of the row-stochastic G = W / rowsum (K <= 32) by Grassmann-Taksar-Heyman
// elimination inside ONE warp: lane i keeps row i in registers, the pivot row travels by shuffles.
// No subtractions, so the result is componentwise accurate (~1e-15); ~4k cycles at K = 16
// (measured alternatives: block-wide elimination through memory ~22k, repeated squaring ~22k).
// Grassmann-Taksar-Heyman elimination for K > 32 (one block, matrix in L2): censor states K-1, ..., 1.
// Step n: every warp forms s = sum_{j<n} G[n][j] (same order in all warps), then for its rows
// i = wp, wp + nw, ...: f = G[i][n] / s (kept for the back-substitution), G[i][j] += f G[n][j], j < n.
// The step is bound by the L2 round trip of the rows, so R rows of a warp are in flight together.
// U: 32-column groups held per lane (K <= 32 U on the register path; U = 2 for K <= 64 keeps the step at ~70
// instead of ~250 warp instructions: the predicated-off groups of the U = 8 form still take issue slots)
template <int R, int U>
__device__ __noinline__ void gth_block(const int K, double* __restrict__ G) {
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5, nw = blockDim.x >> 5;
#pragma unroll 1
  for (int n = K - 1; n >= 1; --n) {
    const double* rown = G + (size_t)n * K;
    double s = 0.0;
    if (K <= 32 * U) {
      double rn[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { const int j = lane + 32 * u; rn[u] = j < n ? rown[j] : 0.0; }
#pragma unroll
      for (int u = 0; u < U; ++u) if (lane + 32 * u < n) s += rn[u];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const double rinv = 1.0 / s;
#pragma unroll 1
      for (int i0 = wp; i0 < n; i0 += R * nw) {
        double* rp[R];
        double g[R], v[R][U];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int i = i0 + r * nw;
          rp[r] = G + (size_t)(i < n ? i : i0) * K;      // rows beyond n: re-read row i0, never stored
          g[r] = rp[r][n];
#pragma unroll
          for (int u = 0; u < U; ++u) { const int j = lane + 32 * u; v[r][u] = j < n ? rp[r][j] : 0.0; }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          g[r] *= rinv;
          if (i0 + r * nw < n) {
#pragma unroll
            for (int u = 0; u < U; ++u) { const int j = lane + 32 * u; if (j < n) rp[r][j] = fma(g[r], rn[u], v[r][u]); }
          }
        }
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int r = 0; r < R; ++r) if (i0 + r * nw < n) rp[r][n] = g[r];
        }
      }
    } else {
#pragma unroll 1
      for (int j = lane; j < n; j += 32) s += rown[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const double rinv = 1.0 / s;
#pragma unroll 1
      for (int i = wp; i < n; i += nw) {
        double* rowi = G + (size_t)i * K;
        const double f = rowi[n] * rinv;
#pragma unroll 4
        for (int j = lane; j < n; j += 32) rowi[j] = fma(f, rown[j], rowi[j]);
        __syncwarp();
        if (lane == 0) rowi[n] = f;
      }
    }
    __syncthreads();
  }
}

template <int KP>
__device__ void gth_warp(const int K, const double* __restrict__ W, const double* __restrict__ rowsum,
                         double* __restrict__ pi_out, const int lane) {
  __shared__ double Gs[KP][KP + 1];                     // eliminated matrix for the back-substitution
  double row[KP];
  const double rs = lane < K ? gth_rcp(rowsum[lane]) : 0.0;
#pragma unroll
  for (int j = 0; j < KP; ++j) row[j] = (lane < K && j < K) ? W[lane * K + j] * rs : 0.0;
#pragma unroll
  for (int n = KP - 1; n >= 1; --n) {
    if (n < K) {                                       // warp-uniform
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;   // four partial sums: the dependent chain is n/4 adds
#pragma unroll
      for (int j = 0; j < n; ++j) {
        if ((j & 3) == 0) s0 += row[j]; else if ((j & 3) == 1) s1 += row[j]; else if ((j & 3) == 2) s2 += row[j]; else s3 += row[j];
      }
      const double s = __shfl_sync(0xffffffffu, (s0 + s1) + (s2 + s3), n);
      const double f = row[n] * gth_rcp(s);
#pragma unroll
      for (int j = 0; j < n; ++j) {
        const double g = __shfl_sync(0xffffffffu, row[j], n);
        if (lane < n) row[j] = fma(f, g, row[j]);
      }
      if (lane < n) row[n] = f;
    }
  }
  // pi[0] = 1; pi[j] = sum_{i<j} pi[i] G[i][j].  G[i][j] (i < j) sits in register row[j] of lane i: transpose
  // through shared memory so that lane j owns column j, then one broadcast + one FMA per i (no reductions)
  if (lane < KP) {
#pragma unroll
    for (int j = 0; j < KP; ++j) Gs[lane][j] = row[j];
  }
  __syncwarp();
  double pv = lane == 0 ? 1.0 : 0.0;
#pragma unroll
  for (int i = 0; i < KP - 1; ++i) {
    if (i + 1 < K) {
      const double pi_i = __shfl_sync(0xffffffffu, pv, i);
      if (lane > i && lane < K) pv = fma(pi_i, Gs[i][lane], pv);
    }
  }
  if (lane < K) pi_out[lane] = pv;
}

// Normalisation of the stationary vector and pi0 = exp(psi(v) - psi(sum v)) (hmmsgd_metaobs.py:418,502),
// one warp.  (Tried: a spare warp executing this code on scratch data beforehand to warm the
// instruction cache -- no change in the clock64 stamps, so the serial tail is not fetch-bound.)
__device__ __noinline__ void pi0_section(const int user_init, const int K, const double* pi, double* vinit,
                                         float* pi0, const int lane, long long* dbg) {
#define PSTAMP(i) do { if (dbg) dbg[i] = clock64(); } while (0)
  PSTAMP(11);
  if (!user_init) {
    double n2 = 0.0;
#pragma unroll 1
    for (int j = lane; j < K; j += 32) n2 += pi[j] * pi[j];
    PSTAMP(12);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
    PSTAMP(13);
    n2 = sqrt(n2);
    PSTAMP(6);
#pragma unroll 1
    for (int j = lane; j < K; j += 32) vinit[j] = fabs(pi[j]) / n2;
  } else {
#pragma unroll 1
    for (int j = lane; j < K; j += 32) vinit[j] = vinit[K + j];
  }
  __syncwarp();
  PSTAMP(7);
  double n1 = 0.0;
#pragma unroll 1
  for (int j = lane; j < K; j += 32) n1 += vinit[j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n1 += __shfl_xor_sync(0xffffffffu, n1, o);
  PSTAMP(8);
  if (K < 32) {                                        // one call: lanes < K their own entry, lane 31 the sum
    const double arg = lane < K ? vinit[lane] : n1;
    const double dg = digamma_fast(arg + SVIHMM_EPS);
    const double dgs = __shfl_sync(0xffffffffu, dg, 31);
    if (lane < K) pi0[lane] = (float)dexp_ni(dg - dgs);
  } else {
    const double dgs = digamma_fast(n1 + SVIHMM_EPS);
#pragma unroll 1
    for (int j = lane; j < K; j += 32) pi0[j] = (float)dexp_ni(digamma_fast(vinit[j] + SVIHMM_EPS) - dgs);
  }
  PSTAMP(9);
#undef PSTAMP
}

__device__ void global_tran_block(const GlobalArgs& a, double* sm) {
  const int K = a.K, tid = threadIdx.x, nth = blockDim.x;
#define GSTAMP(i) do { if (a.dbg && tid == 0) a.dbg[i] = clock64(); } while (0)
#define GSYNC() __syncthreads()
  GSTAMP(0);
  const int KK = K * K;
  const GStats sv = gstats(a.stats, K, a.KE, a.D, a.DD);
  if (a.mode == GM_SVI || a.mode == GM_BSGD) {
#pragma unroll 1
    for (int i = tid; i < KK; i += nth) {
      const double no = a.W[i] - 1.0;
      if (a.ada_G && a.mode == GM_SVI) {                       // :1036-1040: lrate is not used in this branch
        const double G = a.ada_G[i] + no * no;
        a.ada_G[i] = G;
        const double m = sqrt(sqrt(G));
        a.W[i] = (1.0 - 1.0 / m) * no + a.bA * sv.A[i] / m + 1.0;
      } else a.W[i] = (1.0 - a.lrate) * no + a.lrate * a.bA * sv.A[i] + 1.0;
    }
    if (a.mode == GM_BSGD) {
#pragma unroll 1
      for (int i = tid; i < K; i += nth) a.vinit[K + i] = a.prior_init[i] + sv.q0[i];   // hmmbatchsgd.py:216
    }
  } else if (a.mode == GM_BATCH) {
#pragma unroll 1
    for (int i = tid; i < KK; i += nth) a.W[i] = a.prior_tran[i] + sv.A[i];
#pragma unroll 1
    for (int i = tid; i < K; i += nth) a.vinit[K + i] = a.prior_init[i] + sv.q0[i];
  }
  GSYNC();
  // row sums: one warp per row
  const int lane = tid & 31, wp = tid >> 5, nw = nth >> 5;
#pragma unroll 1
  for (int i = wp; i < K; i += nw) {
    double s = 0.0;
#pragma unroll 1
    for (int j = lane; j < K; j += 32) s += a.W[i * K + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) a.rowsum[i] = s;
  }
  GSYNC();
  GSTAMP(1);
  double* pi = sm;                                        // K + 1 doubles
  // K*K doubles of scratch (K > 32 only): in shared memory up to K = 64 (32 KB; a censoring step is then a
  // shared-memory round trip instead of an L2 round trip per row pass), in global memory (L2) beyond
  double* G = (K > 32 && K <= GTH_SMEM_KMAX) ? sm + 2 * K + 2 : a.gth;
  const bool inwarp = !a.user_init && K <= 32;
  // the LAST warp of the block computes the stationary vector while the others do the digamma
  // transforms of the transition matrix
  if (inwarp && wp == nw - 1) {
    if (K <= 2) gth_warp<2>(K, a.W, a.rowsum, pi, lane);
    else if (K <= 4) gth_warp<4>(K, a.W, a.rowsum, pi, lane);
    else if (K <= 8) gth_warp<8>(K, a.W, a.rowsum, pi, lane);
    else if (K <= 16) gth_warp<16>(K, a.W, a.rowsum, pi, lane);
    else gth_warp<32>(K, a.W, a.rowsum, pi, lane);
    if (a.dbg && lane == 0) a.dbg[14] = clock64();
  } else {
    const int nt = inwarp ? nth - 32 : nth;
#pragma unroll 1
    for (int idx = tid; idx < KK; idx += nt) {
      const int i = idx / K, j = idx - i * K;
      const double w = a.W[idx];
      if (!inwarp) G[idx] = w / a.rowsum[i];
      if (a.gth_ext) continue;                               // k_gth_cluster does the transforms on 8 CTAs
      const float v = (float)dexp_ni(digamma_fast(w + SVIHMM_EPS) - digamma_fast(a.rowsum[i] + SVIHMM_EPS));
      a.Pt[idx] = v;
      a.PtT[j * K + i] = v;
    }
  }
  GSYNC();
  GSTAMP(2);
  if (!a.user_init && !inwarp && !a.gth_ext) {
    // Grassmann-Taksar-Heyman: censor states K-1, K-2, ..., 1 (no subtractions)
    if (K <= 64) gth_block<4, 2>(K, G);
    else gth_block<4, 8>(K, G);
    // pi[0] = 1; pi[j] = sum_{i<j} pi[i] G[i][j]: column j accumulates as the pi[i] become final
    if (wp == 0) {
#pragma unroll 1
      for (int j = lane; j < K; j += 32) pi[j] = j == 0 ? 1.0 : 0.0;
      __syncwarp();
#pragma unroll 1
      for (int i = 0; i < K - 1; ++i) {
        const double pv = pi[i];
#pragma unroll 1
        for (int j = i + 1 + lane; j < K; j += 32) pi[j] = fma(pv, G[i * K + j], pi[j]);
        __syncwarp();
      }
    }
  }
  if (wp == 0 && (a.user_init || inwarp || !a.gth_ext))
    pi0_section(a.user_init, K, pi, a.vinit, a.pi0, lane, a.dbg && tid == 0 ? a.dbg : nullptr);
  GSTAMP(3);
}

// ---- emission blocks -----------------------------------------------------------------------
// full covariance: one block per state.  sm: 2*D*D + 3*D doubles.
__device__ void global_emit_full_block(const GlobalArgs& a, const int k, double* sm) {
  const int K = a.K, D = a.D, tid = threadIdx.x, nth = blockDim.x;
  double* L = sm; double* Ri = sm + D * D; double* mu_o = Ri + D * D; double* mu_p = mu_o + D; double* mu_n = mu_p + D;
  const GStats sv = gstats(a.stats, K, a.KE, D, a.DD);
  double* p = a.emit + (size_t)k * a.plen;
  const double* pr = a.prior_emit + (size_t)k * a.plen;
  const size_t oS = D, oK = (size_t)D + (size_t)D * D, oN = oK + 1;
  if (a.mode == GM_SVI || a.mode == GM_BSGD) {
    const double ka_o = p[oK], nu_o = p[oN], ka_p = pr[oK], nu_p = pr[oN], nk = sv.n[k];
    const double e2 = (1.0 - a.lrate) * ka_o + a.lrate * (ka_p + a.bE * nk);
    const double e4 = (1.0 - a.lrate) * (nu_o + 2.0 + D) + a.lrate * (nu_p + 2.0 + D + a.bE * nk);
#pragma unroll 1
    for (int d = tid; d < D; d += nth) {
      mu_o[d] = p[d]; mu_p[d] = pr[d];
      const double e1 = (1.0 - a.lrate) * ka_o * p[d] + a.lrate * (ka_p * pr[d] + a.bE * sv.sx[(size_t)k * D + d]);
      mu_n[d] = e1 / e2;
    }
    __syncthreads();
#pragma unroll 1
    for (int idx = tid; idx < D * D; idx += nth) {
      const int d1 = idx / D, d2 = idx - d1 * D;
      const double e3 = (1.0 - a.lrate) * (p[oS + idx] + ka_o * mu_o[d1] * mu_o[d2])
                      + a.lrate * (pr[oS + idx] + ka_p * mu_p[d1] * mu_p[d2] + a.bE * sv.sxx[(size_t)k * D * D + idx]);
      p[oS + idx] = e3 - mu_n[d1] * mu_n[d2] * e2;
    }
    __syncthreads();
#pragma unroll 1
    for (int d = tid; d < D; d += nth) p[d] = mu_n[d];
    if (tid == 0) { p[oK] = e2; p[oN] = e4 - 2.0 - D; }
    __syncthreads();
  } else if (a.mode == GM_BATCH) {
    const double n = sv.n[k], ka0 = pr[oK], nu0 = pr[oN];
    if (!(n > SVIHMM_WEPS)) {                              // distributions.py:267,275-276: keep the prior
#pragma unroll 1
      for (int idx = tid; idx < (int)a.plen; idx += nth) p[idx] = pr[idx];
    } else {
#pragma unroll 1
      for (int d = tid; d < D; d += nth) mu_o[d] = sv.sx[(size_t)k * D + d] / n;   // xbar
      __syncthreads();
#pragma unroll 1
      for (int idx = tid; idx < D * D; idx += nth) {
        const int d1 = idx / D, d2 = idx - d1 * D;
        const double sumsq = sv.sxx[(size_t)k * D * D + idx] - n * mu_o[d1] * mu_o[d2];
        p[oS + idx] = pr[oS + idx] + sumsq + ka0 * n / (ka0 + n) * (mu_o[d1] - pr[d1]) * (mu_o[d2] - pr[d2]);
      }
#pragma unroll 1
      for (int d = tid; d < D; d += nth) p[d] = ka0 / (ka0 + n) * pr[d] + n / (ka0 + n) * mu_o[d];
      if (tid == 0) { p[oK] = ka0 + n; p[oN] = nu0 + n; }
    }
    __syncthreads();
  }
  // constants: Rs = sqrt(nu/2) chol(sigma)^-1 (packed lower), gk = Rs mu, ck  so that ll = ck - |Rs x - gk|^2
  const double kappa = p[oK], nu = p[oN];
#pragma unroll 1
  for (int idx = tid; idx < D * D; idx += nth) { L[idx] = p[oS + idx]; Ri[idx] = 0.0; }
  __syncthreads();
#pragma unroll 1
  for (int j = 0; j < D; ++j) {
    if (tid == 0) {
      double s = L[j * D + j];
#pragma unroll 1
      for (int q = 0; q < j; ++q) s -= L[j * D + q] * L[j * D + q];
      // sigma = eta3 - kappa mu mu^T cancels when |mean| >> spread and the float32 statistics lose the
      // variance: flag it instead of propagating NaN silently (svihmm_check)
      if (!(s > 0.0)) { atomicOr(a.status, 1); s = 1e-300; }
      L[j * D + j] = sqrt(s);
    }
    __syncthreads();
    const double djj = L[j * D + j];
#pragma unroll 1
    for (int i = j + 1 + tid; i < D; i += nth) {
      double s = L[i * D + j];
#pragma unroll 1
      for (int q = 0; q < j; ++q) s -= L[i * D + q] * L[j * D + q];
      L[i * D + j] = s / djj;
    }
    __syncthreads();
  }
#pragma unroll 1
  for (int c = tid; c < D; c += nth) {                     // column c of L^-1 by forward substitution
    Ri[c * D + c] = 1.0 / L[c * D + c];
#pragma unroll 1
    for (int i = c + 1; i < D; ++i) {
      double s = 0.0;
#pragma unroll 1
      for (int q = c; q < i; ++q) s += L[i * D + q] * Ri[q * D + c];
      Ri[i * D + c] = -s / L[i * D + i];
    }
  }
  __syncthreads();
  const double sc = sqrt(0.5 * nu);
  const size_t tri = (size_t)D * (D + 1) / 2;
#pragma unroll 1
  for (int idx = tid; idx < D * D; idx += nth) {
    const int i = idx / D, j = idx - i * D;
    if (j <= i) a.Rs[k * tri + (size_t)i * (i + 1) / 2 + j] = sc * Ri[idx];
  }
#pragma unroll 1
  for (int i = tid; i < D; i += nth) {
    double s = 0.0;
#pragma unroll 1
    for (int j = 0; j <= i; ++j) s += sc * Ri[i * D + j] * p[j];
    a.gk[(size_t)k * D + i] = s;
  }
  if (tid == 0) {
    double ld = 0.0, dg = 0.0;
#pragma unroll 1
    for (int d = 0; d < D; ++d) { ld += dlog_ni(L[d * D + d]); dg += digamma_fast(0.5 * (nu - d)); }
    a.ck[k] = 0.5 * (dg + D * M_LN2 - 2.0 * ld) - D / (2.0 * kappa) - 0.5 * D * 1.8378770664093453;
  }
}

// diagonal (D independent 1-D NIWs per state): one thread per (k, d), all states in one block sweep
__device__ void global_emit_diag_block(const GlobalArgs& a, const int blk, const int nblk) {
  const int K = a.K, D = a.D, tid = threadIdx.x, nth = blockDim.x;
  const GStats sv = gstats(a.stats, K, a.KE, D, a.DD);
  // block b owns states [k0, k1): whole states per block so that ck can be summed locally
  const int KE = a.KE;
  const int per = (KE + nblk - 1) / nblk, k0 = blk * per, k1 = min(KE, k0 + per);
#pragma unroll 1
  for (int e = k0 * D + tid; e < k1 * D; e += nth) {
    const int k = e / D, d = e - k * D;
    double* p = a.emit + (size_t)k * 4 * D;
    const double* pr = a.prior_emit + (size_t)k * 4 * D;
    double mu = p[d], sg = p[D + d], ka = p[2 * D + d], nu = p[3 * D + d];
    const double mu0 = pr[d], sg0 = pr[D + d], ka0 = pr[2 * D + d], nu0 = pr[3 * D + d];
    if (a.mode == GM_SVI || a.mode == GM_BSGD) {
      const double nk = sv.n[k];
      const double e1 = (1.0 - a.lrate) * ka * mu + a.lrate * (ka0 * mu0 + a.bE * sv.sx[e]);
      const double e2 = (1.0 - a.lrate) * ka + a.lrate * (ka0 + a.bE * nk);
      const double e3 = (1.0 - a.lrate) * (sg + ka * mu * mu) + a.lrate * (sg0 + ka0 * mu0 * mu0 + a.bE * sv.sxx[e]);
      const double e4 = (1.0 - a.lrate) * (nu + 3.0) + a.lrate * (nu0 + 3.0 + a.bE * nk);
      mu = e1 / e2; sg = e3 - mu * mu * e2; ka = e2; nu = e4 - 3.0;
    } else if (a.mode == GM_BATCH) {
      const double n = sv.n[k];
      if (!(n > SVIHMM_WEPS)) { mu = mu0; sg = sg0; ka = ka0; nu = nu0; }
      else {
        const double xb = sv.sx[e] / n, sumsq = sv.sxx[e] - n * xb * xb;
        mu = ka0 / (ka0 + n) * mu0 + n / (ka0 + n) * xb;
        sg = sg0 + sumsq + ka0 * n / (ka0 + n) * (xb - mu0) * (xb - mu0);
        ka = ka0 + n; nu = nu0 + n;
      }
    }
    if (a.mode != GM_PREP) { p[d] = mu; p[D + d] = sg; p[2 * D + d] = ka; p[3 * D + d] = nu; }
    if (!(sg > 0.0)) atomicOr(a.status, 1);                // variance lost to cancellation (see svihmm_check)
    // ll = ck - sum_d Rs (x_d - mu_d)^2 with Rs = nu / (2 sigma)
    const double rs = nu / (2.0 * sg);
    a.Rs[e] = rs;
    a.gk[e] = mu;
    a.par2[2 * ((size_t)d * KE + k)] = -rs;
    a.par2[2 * ((size_t)d * KE + k) + 1] = 2.0 * rs * mu;
    a.ckc[2 * (size_t)e] = 0.5 * (digamma_fast(0.5 * nu) + M_LN2 - dlog_ni(sg)) - 1.0 / (2.0 * ka) - 0.5 * 1.8378770664093453;
    a.ckc[2 * (size_t)e + 1] = rs * mu * mu;
  }
  __syncthreads();
#pragma unroll 1
  for (int k = k0 + tid; k < k1; k += nth) {
    double c = 0.0, c0 = 0.0;
#pragma unroll 1
    for (int d = 0; d < D; ++d) { c += a.ckc[2 * ((size_t)k * D + d)]; c0 += a.ckc[2 * ((size_t)k * D + d) + 1]; }
    a.ck[k] = c;
    a.ckp[k] = c - c0;
  }
}

// categorical (Dirichlet over D symbols per state): one block per state.  sm: 1 double.
//   SVI   hmmsgd_metaobs.py:1071-1084 with emit_inter = sum over the minibatch's windows of
//         (alphav_0 + counts - 1) (:925-926): alpha <- (1-rho)(alpha-1) + rho bE (B (alpha0-1) + counts) + 1
//   BATCH alpha = alphav_0 + counts (Categorical.meanfieldupdate, distributions.py:1366-1370)
//   BSGD  alpha <- (1-rho)(alpha-1) + rho (alphav_0 + counts - 1) + 1 (hmmbatchsgd.py:202-259 pattern)
// then logp[k][c] = psi(alpha[c]) - psi(sum_c alpha[c]) (distributions.py:1383-1386) into Rs.
__device__ void global_emit_cat_block(const GlobalArgs& a, const int k, double* sm) {
  const int K = a.K, C = a.D, tid = threadIdx.x, nth = blockDim.x;
  const GStats sv = gstats(a.stats, K, K, C, 0);
  double* p = a.emit + (size_t)k * C;
  const double* pr = a.prior_emit + (size_t)k * C;
  if (a.mode != GM_PREP) {
    const double nwin = a.mode == GM_SVI ? sv.q0[K + 2] : 1.0;       // statistics tail [.., .., B, ..]
#pragma unroll 1
    for (int c = tid; c < C; c += nth) {
      const double cnt = sv.sx[(size_t)k * C + c];
      double v;
      if (a.mode == GM_SVI) v = (1.0 - a.lrate) * (p[c] - 1.0) + a.lrate * a.bE * (nwin * (pr[c] - 1.0) + cnt) + 1.0;
      else if (a.mode == GM_BSGD) v = (1.0 - a.lrate) * (p[c] - 1.0) + a.lrate * (pr[c] + cnt - 1.0) + 1.0;
      else v = pr[c] + cnt;
      p[c] = v;
    }
    __syncthreads();
  }
  if (tid < 32) {
    double s = 0.0;
#pragma unroll 1
    for (int c = tid; c < C; c += 32) s += p[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (tid == 0) sm[0] = digamma_d(s);
  }
  __syncthreads();
  const double dgs = sm[0];
#pragma unroll 1
  for (int c = tid; c < C; c += nth) a.Rs[(size_t)k * C + c] = digamma_d(p[c]) - dgs;
}

// mixture weights (EXTENSION, BASELINE config 5): Dirichlet(omega_k) over the C components of state k.
//   SVI  omega <- (1-rho)(omega-1) + rho (omega0 - 1 + bE n_kc) + 1   (the transition-row rule
//        hmmsgd_metaobs.py:1029-1045 applied to the component counts)
// then lw[kc] = psi(omega_kc) - psi(sum_c omega_kc) (pybasicbayes/internals/labels.py:59,
// distributions.py:1383-1386).
__device__ void global_mix_block(const GlobalArgs& a) {
  const int K = a.K, C = a.C, KE = a.KE, tid = threadIdx.x, nth = blockDim.x;
  const GStats sv = gstats(a.stats, K, KE, a.D, a.DD);
  if (a.mode == GM_SVI) {
#pragma unroll 1
    for (int e = tid; e < KE; e += nth)
      a.omega[e] = (1.0 - a.lrate) * (a.omega[e] - 1.0) + a.lrate * (a.omega_prior[e] - 1.0 + a.bE * sv.n[e]) + 1.0;
    __syncthreads();
  }
#pragma unroll 1
  for (int k = tid; k < K; k += nth) {
    double s = 0.0;
#pragma unroll 1
    for (int c = 0; c < C; ++c) s += a.omega[k * C + c];
    const double dgs = digamma_d(s);
#pragma unroll 1
    for (int c = 0; c < C; ++c) a.lw[k * C + c] = digamma_d(a.omega[k * C + c]) - dgs;
  }
}

// grid: 1 + (diag ? nblk_diag : K) blocks.
__global__ void __launch_bounds__(512) k_global_step(const GlobalArgs a, const int nblk_emit) {
  extern __shared__ double gsm[];
  // programmatic dependent launch (svihmm_svi_run): let the next E-step's CTAs be scheduled as soon as
  // there is room, then wait for the E-step whose statistics this update consumes (both are no-ops for a
  // plain stream-ordered launch)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (a.zero_buf && blockIdx.x == 0)
    for (size_t i = threadIdx.x; i < a.slen; i += blockDim.x) a.zero_buf[i] = 0.0;
  if (a.world > 1 && a.mode != GM_PREP) {
    const size_t KK = (size_t)a.K * a.K, o_n = KK, o_sx = o_n + a.KE, o_sxx = o_sx + (size_t)a.KE * a.D,
                 o_q0 = o_sxx + (size_t)a.KE * a.DD;
    CommRange rg[4];
    int nrg = 0;
    rg[nrg++] = {o_q0 + a.K, a.slen};                            // tail: every block (B count, bounds)
    if (blockIdx.x == 0) { rg[nrg++] = {0, KK}; rg[nrg++] = {o_q0, o_q0 + a.K}; }
    else if ((int)blockIdx.x == 1 + nblk_emit) rg[nrg++] = {o_n, o_sx};
    else {
      int k0 = blockIdx.x - 1, k1 = k0 + 1;                      // full / categorical: one component per block
      if (a.diag) { const int per = (a.KE + nblk_emit - 1) / nblk_emit; k0 = (blockIdx.x - 1) * per; k1 = min(a.KE, k0 + per); }
      if (k0 < k1) {
        rg[nrg++] = {o_n + k0, o_n + k1};
        rg[nrg++] = {o_sx + (size_t)k0 * a.D, o_sx + (size_t)k1 * a.D};
        rg[nrg++] = {o_sxx + (size_t)k0 * a.DD, o_sxx + (size_t)k1 * a.DD};
      }
    }
    comm_exchange(a, rg, nrg);
    __threadfence();
    __syncthreads();
  }
  if (blockIdx.x == 0) global_tran_block(a, gsm);
  else if ((int)blockIdx.x == 1 + nblk_emit) global_mix_block(a);
  else if (a.cat) global_emit_cat_block(a, blockIdx.x - 1, gsm);
  else if (a.diag) {
    if (a.dbg && threadIdx.x == 0) a.dbg[4] = clock64();
    global_emit_diag_block(a, blockIdx.x - 1, nblk_emit);
    if (a.dbg && threadIdx.x == 0) a.dbg[5] = clock64();
  }
  else global_emit_full_block(a, blockIdx.x - 1, gsm);
}
