// Global part of the variational lower bound on the device:
//   hmmsgd_metaobs.VBHMM.global_lower_bound (hmmsgd_metaobs.py:273-296) = Dirichlet energy + entropy of the
//   transition rows + sum_k var_emit[k].get_vlb(); hmmbase.lower_bound (hmmbase.py:145-199) adds the
//   Dirichlet terms of the initial distribution.
//   Gaussian.get_vlb          pybasicbayes/distributions.py:331-349 (Bishop 10.74 / 10.77); the inverse-Wishart
//                             entropy and log partition function live in the absent pymattutil and follow
//                             their textbook form (as the host classes do: parity unpinned for those two)
//   Categorical.get_vlb       distributions.py:1372-1381
// One block per term: block 0 the transition (and initial) Dirichlets, blocks 1..KE the emission
// components, the last block the mixture weights; every block adds its float64 term to out[0] (zeroed by
// the caller).  All sums are thread-serial or fixed-order tree sums inside a block; the cross-block sum
// is a float64 atomic (KE + 2 terms).
#pragma once
#include "global.cuh"

struct BoundArgs {
  int K, D, KE, C, kind, include_init;
  size_t plen;
  const double *W, *vinit, *emit, *prior_tran, *prior_init, *prior_emit, *omega, *omega_prior;
  double* out;
};

__device__ __forceinline__ double bd_block_sum(double v, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  return t;
}
// energy + entropy of q = Dir(qv) under the prior Dir(pv), n entries (hmmsgd_metaobs.py:277-288 with eps)
__device__ double bd_dirichlet(const double* pv, const double* qv, const int n, double* red, const double eps) {
  double sp = 0.0, sq = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) { sp += pv[i]; sq += qv[i]; }
  sp = bd_block_sum(sp, red); sq = bd_block_sum(sq, red);
  const double dgs = digamma_d(sq + eps);
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double e = digamma_d(qv[i] + eps) - dgs;
    acc += -lgamma(pv[i] + eps) + (pv[i] - 1.0) * e + lgamma(qv[i] + eps) - (qv[i] - 1.0) * e;
  }
  acc = bd_block_sum(acc, red);
  return acc + lgamma(sp + eps) - lgamma(sq + eps);
}

// in-place Cholesky of the D x D matrix in shared memory (lower triangle), as in global_emit_full_block
__device__ void bd_chol(double* L, const int D) {
  const int tid = threadIdx.x, nth = blockDim.x;
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
}
// log partition function of the inverse Wishart (sum of log diag chol = ld)
__device__ double bd_iw_logz(const double ld, const double nu, const int D) {
  double g = 0.0;
  for (int d = 0; d < D; ++d) g += lgamma(0.5 * (nu - d));
  return -(nu * ld - (0.5 * nu * D * M_LN2 + 0.25 * D * (D - 1) * 1.1447298858494002 + g));
}

__global__ void __launch_bounds__(256) k_global_bound(const BoundArgs a) {
  extern __shared__ double bsm[];                       // full NIW: 2*D*D + 2*D doubles
  __shared__ double red[8];
  const int K = a.K, D = a.D, tid = threadIdx.x, nth = blockDim.x;
  const double eps = SVIHMM_EPS;
  double term = 0.0;
  if (blockIdx.x == 0) {
    for (int i = 0; i < K; ++i) term += bd_dirichlet(a.prior_tran + (size_t)i * K, a.W + (size_t)i * K, K, red, eps);
    if (a.include_init) term += bd_dirichlet(a.prior_init, a.vinit, K, red, eps);
  } else if (blockIdx.x <= a.KE) {
    const int k = blockIdx.x - 1;
    const double* p = a.emit + (size_t)k * a.plen;
    const double* pr = a.prior_emit + (size_t)k * a.plen;
    if (a.kind == SVIHMM_EMIT_CATEGORICAL) {
      // Categorical.get_vlb has no eps (distributions.py:1372-1381)
      term = bd_dirichlet(pr, p, D, red, 0.0);
    } else if (a.kind == SVIHMM_EMIT_NIW_DIAG) {
      // D independent one-dimensional NIW factors: Gaussian.get_vlb with D = 1 per dimension
      double acc = 0.0;
      for (int d = tid; d < D; d += nth) {
        const double mu = p[d], sg = p[D + d], ka = p[2 * D + d], nu = p[3 * D + d];
        const double mu0 = pr[d], sg0 = pr[D + d], ka0 = pr[2 * D + d], nu0 = pr[3 * D + d];
        const double llt = digamma_d(0.5 * nu) + M_LN2 - log(sg);
        const double lz = bd_iw_logz(0.5 * log(sg), nu, 1), lz0 = bd_iw_logz(0.5 * log(sg0), nu0, 1);
        const double ent = lz - 0.5 * (nu - 2.0) * llt + 0.5 * nu;
        const double q_entropy = -0.5 * (llt + (log(ka / (2.0 * M_PI)) - 1.0)) + ent;
        const double dm = mu - mu0;
        const double p_avg = 0.5 * (log(ka0 / (2.0 * M_PI)) + llt - ka0 / ka - ka0 * nu * dm * dm / sg)
                           + lz0 + 0.5 * (nu0 - 2.0) * llt - 0.5 * nu * sg0 / sg;
        acc += p_avg + q_entropy;
      }
      term = bd_block_sum(acc, red);
    } else {
      double* L = bsm; double* L0 = bsm + D * D; double* y = L0 + D * D; double* dmu = y + D;
      const size_t oS = D, oK = (size_t)D + (size_t)D * D, oN = oK + 1;
      const double ka = p[oK], nu = p[oN], ka0 = pr[oK], nu0 = pr[oN];
      for (int i = tid; i < D * D; i += nth) { L[i] = p[oS + i]; L0[i] = pr[oS + i]; }
      for (int d = tid; d < D; d += nth) dmu[d] = p[d] - pr[d];
      __syncthreads();
      bd_chol(L, D); bd_chol(L0, D);
      // tr(Sigma_mf^-1 Sigma_0) = |L^-1 L0|_F^2: column c of L^-1 L0 by forward substitution, thread per column;
      // the quadratic form dmu^T Sigma_mf^-1 dmu = |L^-1 dmu|^2 is "column D"
      double part = 0.0;
      for (int c = tid; c <= D; c += nth) {
        double* col = nullptr;                          // no scratch per thread: recompute y_i on the fly (D <= 96)
        (void)col;
        double yv[96];
        double ss = 0.0;
        for (int i = 0; i < D; ++i) {
          double s = c < D ? (i >= c ? L0[i * D + c] : 0.0) : dmu[i];
          for (int q = 0; q < i; ++q) s -= L[i * D + q] * yv[q];
          yv[i] = s / L[i * D + i];
          ss += yv[i] * yv[i];
        }
        part += c < D ? -0.5 * nu * ss : -0.5 * ka0 * nu * ss;
      }
      part = bd_block_sum(part, red);
      if (tid == 0) {
        double ld = 0.0, ld0 = 0.0, dg = 0.0;
        for (int d = 0; d < D; ++d) { ld += log(L[d * D + d]); ld0 += log(L0[d * D + d]); dg += digamma_d(0.5 * (nu - d)); }
        const double llt = dg + D * M_LN2 - 2.0 * ld;
        const double ent = bd_iw_logz(ld, nu, D) - 0.5 * (nu - D - 1.0) * llt + 0.5 * nu * D;
        const double q_entropy = -0.5 * (llt + D * (log(ka / (2.0 * M_PI)) - 1.0)) + ent;
        const double p_avg = 0.5 * (D * log(ka0 / (2.0 * M_PI)) + llt - D * ka0 / ka) + bd_iw_logz(ld0, nu0, D)
                           + 0.5 * (nu0 - D - 1.0) * llt;
        term = p_avg + q_entropy + part;
      }
    }
  } else {
    // mixture weights: Categorical.get_vlb of every state's Dirichlet over its C components
    for (int k = 0; k < K; ++k) term += bd_dirichlet(a.omega_prior + (size_t)k * a.C, a.omega + (size_t)k * a.C, a.C, red, 0.0);
  }
  if (tid == 0) atomicAdd(a.out, term);
}
