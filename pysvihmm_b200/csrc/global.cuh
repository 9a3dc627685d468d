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

enum { GM_PREP = 0, GM_SVI = 1, GM_BATCH = 2 };

struct GlobalArgs {
  int K, D, DD, diag, mode, user_init;
  size_t plen;
  double *W, *vinit, *emit;                 // vinit: [0,K) the vector in use, [K,2K) the user-given one
  const double *prior_tran, *prior_init, *prior_emit, *stats;
  double lrate, bA, bE;
  double *gth, *rowsum, *ckc;               // scratch: 2*K*K, K, 2*K*D doubles
  float *Pt, *PtT, *pi0;
  double *Rs, *gk, *ck;
  double *par2, *ckp;                       // diagonal, fused-kernel form: [d][k] (-Rs, 2 Rs mu) and ck - sum Rs mu^2
};

// psi(x), float64: recurrence up to x >= 10 (branch-free, the reciprocals are independent), then the
// asymptotic series through B14 (truncation error < 5e-17 there); reflection for x <= 0.
__device__ inline double digamma_fast(double x) {
  double r = 0.0;
  if (x <= 0.0) {
    if (x == floor(x)) return nan("");
    r = -M_PI / tan(M_PI * x);
    x = 1.0 - x;
  }
  if (x < 10.0) {
    const int n = (int)ceil(10.0 - x);               // 1..10 shifts
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int i = 0; i < 10; i += 2) {
      if (i < n) s0 += 1.0 / (x + i);
      if (i + 1 < n) s1 += 1.0 / (x + (i + 1));
    }
    r -= s0 + s1;
    x += n;
  }
  const double xi = 1.0 / x, x2 = xi * xi;
  r += log(x) - 0.5 * xi
     - x2 * (1.0 / 12 - x2 * (1.0 / 120 - x2 * (1.0 / 252 - x2 * (1.0 / 240
     - x2 * (1.0 / 132 - x2 * (691.0 / 32760 - x2 * (1.0 / 12)))))));
  return r;
}

struct GStats { const double *A, *n, *sx, *sxx, *q0; };
__device__ inline GStats gstats(const double* s, int K, int D, int DD) {
  GStats v;
  v.A = s; v.n = s + (size_t)K * K; v.sx = v.n + K; v.sxx = v.sx + (size_t)K * D; v.q0 = v.sxx + (size_t)K * DD;
  return v;
}

// ---- block 0 -------------------------------------------------------------------------------
__device__ __forceinline__ void bar_named(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Perron vector of the row-stochastic G (K <= 32) by repeated squaring: G^(2^m) -> 1 pi^T at the
// rate lambda_2^(2^m), so ~log2(log(1e-16)/log(lambda_2)) squarings of a K x K matrix (each a
// fully parallel K^3 product) replace the K-step elimination.  All entries of G are positive
// (Dirichlet parameters > 0), hence the chain is primitive.  Executed by threads [t0, t0+nt).
__device__ void stationary_by_squaring(const int K, double* A, double* Bm, double* pi, const int tl, const int nt,
                                       const int bar_id) {
  const int KK = K * K;
  for (int it = 0; it < 60; ++it) {
    for (int idx = tl; idx < KK; idx += nt) {
      const int i = idx / K, j = idx - i * K;
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int l = 0;
      for (; l + 3 < K; l += 4) {
        s0 = fma(A[i * K + l], A[l * K + j], s0);
        s1 = fma(A[i * K + l + 1], A[(l + 1) * K + j], s1);
        s2 = fma(A[i * K + l + 2], A[(l + 2) * K + j], s2);
        s3 = fma(A[i * K + l + 3], A[(l + 3) * K + j], s3);
      }
      for (; l < K; ++l) s0 = fma(A[i * K + l], A[l * K + j], s0);
      Bm[idx] = (s0 + s1) + (s2 + s3);
    }
    bar_named(bar_id, nt);
    // converged when every row agrees with row 0 to 1e-13 relative (rounding noise is ~K eps)
    int bad = 0;
    for (int idx = tl; idx < KK; idx += nt) {
      const int j = idx % K;
      bad |= fabs(Bm[idx] - Bm[j]) > 1e-13 * Bm[j];
    }
    double* t = A; A = Bm; Bm = t;
    // block-wide OR over the nt participating threads through shared memory
    if (tl == 0) pi[K] = 0.0;
    bar_named(bar_id, nt);
    if (bad) pi[K] = 1.0;
    bar_named(bar_id, nt);
    if (pi[K] == 0.0) break;
  }
  for (int j = tl; j < K; j += nt) pi[j] = A[j];
}

__device__ void global_tran_block(const GlobalArgs& a, double* sm) {
  const int K = a.K, tid = threadIdx.x, nth = blockDim.x;
  const int KK = K * K;
  const GStats sv = gstats(a.stats, K, a.D, a.DD);
  if (a.mode == GM_SVI) {
    for (int i = tid; i < KK; i += nth) a.W[i] = (1.0 - a.lrate) * (a.W[i] - 1.0) + a.lrate * a.bA * sv.A[i] + 1.0;
  } else if (a.mode == GM_BATCH) {
    for (int i = tid; i < KK; i += nth) a.W[i] = a.prior_tran[i] + sv.A[i];
    for (int i = tid; i < K; i += nth) a.vinit[K + i] = a.prior_init[i] + sv.q0[i];
  }
  __syncthreads();
  // row sums: one warp per row
  const int lane = tid & 31, wp = tid >> 5, nw = nth >> 5;
  for (int i = wp; i < K; i += nw) {
    double s = 0.0;
    for (int j = lane; j < K; j += 32) s += a.W[i * K + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) a.rowsum[i] = s;
  }
  __syncthreads();
  double* pi = sm;                                        // K + 1 doubles
  double* G = a.gth;                                      // K*K (+ K*K for the squaring) doubles of scratch
  const bool square = !a.user_init && K <= 32;
  const int half = nth / 2;
  // the two halves of the block work concurrently: threads [0, half) the digamma transforms of the
  // transition matrix, threads [half, nth) the stationary vector
  if (tid < half || !square) {
    const int tl = square ? tid : tid, nt = square ? half : nth;
    for (int idx = tl; idx < KK; idx += nt) {
      const int i = idx / K, j = idx - i * K;
      const double w = a.W[idx];
      const float v = (float)exp(digamma_fast(w + SVIHMM_EPS) - digamma_fast(a.rowsum[i] + SVIHMM_EPS));
      a.Pt[idx] = v;
      a.PtT[j * K + i] = v;
      if (!square) G[idx] = w / a.rowsum[i];
    }
  } else {
    const int tl = tid - half, nt = nth - half;
    for (int idx = tl; idx < KK; idx += nt) G[idx] = a.W[idx] / a.rowsum[idx / K];
    bar_named(1, nt);
    stationary_by_squaring(K, G, G + KK, pi, tl, nt, 1);
  }
  __syncthreads();
  if (!a.user_init && !square) {
    // Grassmann-Taksar-Heyman: censor states K-1, K-2, ..., 1 (no subtractions)
    for (int n = K - 1; n >= 1; --n) {
      if (wp == 0) {
        double s = 0.0;
        for (int j = lane; j < n; j += 32) s += G[n * K + j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const double rinv = 1.0 / s;
        for (int i = lane; i < n; i += 32) G[i * K + n] *= rinv;
      }
      __syncthreads();
      for (int idx = tid; idx < n * n; idx += nth) {
        const int i = idx / n, j = idx - i * n;
        G[i * K + j] = fma(G[i * K + n], G[n * K + j], G[i * K + j]);
      }
      __syncthreads();
    }
    // pi[0] = 1; pi[j] = sum_{i<j} pi[i] G[i][j]: column j accumulates as the pi[i] become final
    if (wp == 0) {
      for (int j = lane; j < K; j += 32) pi[j] = j == 0 ? 1.0 : 0.0;
      __syncwarp();
      for (int i = 0; i < K - 1; ++i) {
        const double pv = pi[i];
        for (int j = i + 1 + lane; j < K; j += 32) pi[j] = fma(pv, G[i * K + j], pi[j]);
        __syncwarp();
      }
    }
  }
  if (wp == 0) {
    if (!a.user_init) {
      double n2 = 0.0;
      for (int j = lane; j < K; j += 32) n2 += pi[j] * pi[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) n2 += __shfl_xor_sync(0xffffffffu, n2, o);
      n2 = sqrt(n2);
      for (int j = lane; j < K; j += 32) a.vinit[j] = fabs(pi[j]) / n2;
    } else {
      for (int j = lane; j < K; j += 32) a.vinit[j] = a.vinit[K + j];
    }
    __syncwarp();
    double n1 = 0.0;
    for (int j = lane; j < K; j += 32) n1 += a.vinit[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n1 += __shfl_xor_sync(0xffffffffu, n1, o);
    const double dgs = digamma_fast(n1 + SVIHMM_EPS);
    for (int j = lane; j < K; j += 32) a.pi0[j] = (float)exp(digamma_fast(a.vinit[j] + SVIHMM_EPS) - dgs);
  }
}

// ---- emission blocks -----------------------------------------------------------------------
// full covariance: one block per state.  sm: 2*D*D + 3*D doubles.
__device__ void global_emit_full_block(const GlobalArgs& a, const int k, double* sm) {
  const int K = a.K, D = a.D, tid = threadIdx.x, nth = blockDim.x;
  double* L = sm; double* Ri = sm + D * D; double* mu_o = Ri + D * D; double* mu_p = mu_o + D; double* mu_n = mu_p + D;
  const GStats sv = gstats(a.stats, K, D, a.DD);
  double* p = a.emit + (size_t)k * a.plen;
  const double* pr = a.prior_emit + (size_t)k * a.plen;
  const size_t oS = D, oK = (size_t)D + (size_t)D * D, oN = oK + 1;
  if (a.mode == GM_SVI) {
    const double ka_o = p[oK], nu_o = p[oN], ka_p = pr[oK], nu_p = pr[oN], nk = sv.n[k];
    const double e2 = (1.0 - a.lrate) * ka_o + a.lrate * (ka_p + a.bE * nk);
    const double e4 = (1.0 - a.lrate) * (nu_o + 2.0 + D) + a.lrate * (nu_p + 2.0 + D + a.bE * nk);
    for (int d = tid; d < D; d += nth) {
      mu_o[d] = p[d]; mu_p[d] = pr[d];
      const double e1 = (1.0 - a.lrate) * ka_o * p[d] + a.lrate * (ka_p * pr[d] + a.bE * sv.sx[(size_t)k * D + d]);
      mu_n[d] = e1 / e2;
    }
    __syncthreads();
    for (int idx = tid; idx < D * D; idx += nth) {
      const int d1 = idx / D, d2 = idx - d1 * D;
      const double e3 = (1.0 - a.lrate) * (p[oS + idx] + ka_o * mu_o[d1] * mu_o[d2])
                      + a.lrate * (pr[oS + idx] + ka_p * mu_p[d1] * mu_p[d2] + a.bE * sv.sxx[(size_t)k * D * D + idx]);
      p[oS + idx] = e3 - mu_n[d1] * mu_n[d2] * e2;
    }
    __syncthreads();
    for (int d = tid; d < D; d += nth) p[d] = mu_n[d];
    if (tid == 0) { p[oK] = e2; p[oN] = e4 - 2.0 - D; }
    __syncthreads();
  } else if (a.mode == GM_BATCH) {
    const double n = sv.n[k], ka0 = pr[oK], nu0 = pr[oN];
    if (!(n > SVIHMM_WEPS)) {                              // distributions.py:267,275-276: keep the prior
      for (int idx = tid; idx < (int)a.plen; idx += nth) p[idx] = pr[idx];
    } else {
      for (int d = tid; d < D; d += nth) mu_o[d] = sv.sx[(size_t)k * D + d] / n;   // xbar
      __syncthreads();
      for (int idx = tid; idx < D * D; idx += nth) {
        const int d1 = idx / D, d2 = idx - d1 * D;
        const double sumsq = sv.sxx[(size_t)k * D * D + idx] - n * mu_o[d1] * mu_o[d2];
        p[oS + idx] = pr[oS + idx] + sumsq + ka0 * n / (ka0 + n) * (mu_o[d1] - pr[d1]) * (mu_o[d2] - pr[d2]);
      }
      for (int d = tid; d < D; d += nth) p[d] = ka0 / (ka0 + n) * pr[d] + n / (ka0 + n) * mu_o[d];
      if (tid == 0) { p[oK] = ka0 + n; p[oN] = nu0 + n; }
    }
    __syncthreads();
  }
  // constants: Rs = sqrt(nu/2) chol(sigma)^-1 (packed lower), gk = Rs mu, ck  so that ll = ck - |Rs x - gk|^2
  const double kappa = p[oK], nu = p[oN];
  for (int idx = tid; idx < D * D; idx += nth) { L[idx] = p[oS + idx]; Ri[idx] = 0.0; }
  __syncthreads();
  for (int j = 0; j < D; ++j) {
    if (tid == 0) {
      double s = L[j * D + j];
      for (int q = 0; q < j; ++q) s -= L[j * D + q] * L[j * D + q];
      L[j * D + j] = sqrt(s);
    }
    __syncthreads();
    const double djj = L[j * D + j];
    for (int i = j + 1 + tid; i < D; i += nth) {
      double s = L[i * D + j];
      for (int q = 0; q < j; ++q) s -= L[i * D + q] * L[j * D + q];
      L[i * D + j] = s / djj;
    }
    __syncthreads();
  }
  for (int c = tid; c < D; c += nth) {                     // column c of L^-1 by forward substitution
    Ri[c * D + c] = 1.0 / L[c * D + c];
    for (int i = c + 1; i < D; ++i) {
      double s = 0.0;
      for (int q = c; q < i; ++q) s += L[i * D + q] * Ri[q * D + c];
      Ri[i * D + c] = -s / L[i * D + i];
    }
  }
  __syncthreads();
  const double sc = sqrt(0.5 * nu);
  const size_t tri = (size_t)D * (D + 1) / 2;
  for (int idx = tid; idx < D * D; idx += nth) {
    const int i = idx / D, j = idx - i * D;
    if (j <= i) a.Rs[k * tri + (size_t)i * (i + 1) / 2 + j] = sc * Ri[idx];
  }
  for (int i = tid; i < D; i += nth) {
    double s = 0.0;
    for (int j = 0; j <= i; ++j) s += sc * Ri[i * D + j] * p[j];
    a.gk[(size_t)k * D + i] = s;
  }
  if (tid == 0) {
    double ld = 0.0, dg = 0.0;
    for (int d = 0; d < D; ++d) { ld += log(L[d * D + d]); dg += digamma_fast(0.5 * (nu - d)); }
    a.ck[k] = 0.5 * (dg + D * M_LN2 - 2.0 * ld) - D / (2.0 * kappa) - 0.5 * D * log(2.0 * M_PI);
  }
}

// diagonal (D independent 1-D NIWs per state): one thread per (k, d), all states in one block sweep
__device__ void global_emit_diag_block(const GlobalArgs& a, const int blk, const int nblk) {
  const int K = a.K, D = a.D, tid = threadIdx.x, nth = blockDim.x;
  const GStats sv = gstats(a.stats, K, D, a.DD);
  // block b owns states [k0, k1): whole states per block so that ck can be summed locally
  const int per = (K + nblk - 1) / nblk, k0 = blk * per, k1 = min(K, k0 + per);
  for (int e = k0 * D + tid; e < k1 * D; e += nth) {
    const int k = e / D, d = e - k * D;
    double* p = a.emit + (size_t)k * 4 * D;
    const double* pr = a.prior_emit + (size_t)k * 4 * D;
    double mu = p[d], sg = p[D + d], ka = p[2 * D + d], nu = p[3 * D + d];
    const double mu0 = pr[d], sg0 = pr[D + d], ka0 = pr[2 * D + d], nu0 = pr[3 * D + d];
    if (a.mode == GM_SVI) {
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
    // ll = ck - sum_d Rs (x_d - mu_d)^2 with Rs = nu / (2 sigma)
    const double rs = nu / (2.0 * sg);
    a.Rs[e] = rs;
    a.gk[e] = mu;
    a.par2[2 * ((size_t)d * K + k)] = -rs;
    a.par2[2 * ((size_t)d * K + k) + 1] = 2.0 * rs * mu;
    a.ckc[2 * (size_t)e] = 0.5 * (digamma_fast(0.5 * nu) + M_LN2 - log(sg)) - 1.0 / (2.0 * ka) - 0.5 * log(2.0 * M_PI);
    a.ckc[2 * (size_t)e + 1] = rs * mu * mu;
  }
  __syncthreads();
  for (int k = k0 + tid; k < k1; k += nth) {
    double c = 0.0, c0 = 0.0;
    for (int d = 0; d < D; ++d) { c += a.ckc[2 * ((size_t)k * D + d)]; c0 += a.ckc[2 * ((size_t)k * D + d) + 1]; }
    a.ck[k] = c;
    a.ckp[k] = c - c0;
  }
}

// grid: 1 + (diag ? nblk_diag : K) blocks.
__global__ void __launch_bounds__(256) k_global_step(const GlobalArgs a, const int nblk_emit) {
  extern __shared__ double gsm[];
  if (blockIdx.x == 0) global_tran_block(a, gsm);
  else if (a.diag) global_emit_diag_block(a, blockIdx.x - 1, nblk_emit);
  else global_emit_full_block(a, blockIdx.x - 1, gsm);
}
