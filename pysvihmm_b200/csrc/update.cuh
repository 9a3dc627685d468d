// K5: global updates on the device-resident float64 master parameters (no host round trip).
//   SVI natural-gradient step   hmmsgd_metaobs.py:1010-1069 with util.py:28-60
//   batch coordinate ascent     hmmbatchcd.py:172-189 with pybasicbayes/distributions.py:240-276
#pragma once
#include "common.cuh"

struct StatsView {           // offsets into the packed statistics (include/svihmm.h)
  const double* A; const double* n; const double* sx; const double* sxx; const double* q0;
};
__host__ __device__ inline StatsView stats_view(const double* s, int K, int D, int DD) {
  StatsView v;
  v.A = s; v.n = s + (size_t)K * K; v.sx = v.n + K; v.sxx = v.sx + (size_t)K * D;
  v.q0 = v.sxx + (size_t)K * DD;
  return v;
}

__global__ void k_update_tran_svi(int KK, double* __restrict__ W, const double* __restrict__ A,
                                  double lrate, double bA) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < KK) W[idx] = (1.0 - lrate) * (W[idx] - 1.0) + lrate * bA * A[idx] + 1.0;
}

// One CTA per state; dynamic smem 3*D doubles (old mu, prior mu, new mu).
__global__ void k_update_emit_svi_full(int K, int D, size_t plen, double* __restrict__ emit,
                                       const double* __restrict__ prior, const double* __restrict__ stats,
                                       double lrate, double bE) {
  extern __shared__ double sm[];
  double* mu_o = sm; double* mu_p = sm + D; double* mu_n = sm + 2 * D;
  const int k = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  const StatsView sv = stats_view(stats, K, D, D * D);
  double* p = emit + (size_t)k * plen;
  const double* pr = prior + (size_t)k * plen;
  const double ka_o = p[D + D * D], nu_o = p[D + D * D + 1];
  const double ka_p = pr[D + D * D], nu_p = pr[D + D * D + 1];
  const double nk = sv.n[k];
  const double e2 = (1.0 - lrate) * ka_o + lrate * (ka_p + bE * nk);
  const double e4 = (1.0 - lrate) * (nu_o + 2.0 + D) + lrate * (nu_p + 2.0 + D + bE * nk);
  for (int d = tid; d < D; d += nth) {
    mu_o[d] = p[d]; mu_p[d] = pr[d];
    const double e1 = (1.0 - lrate) * ka_o * p[d] + lrate * (ka_p * pr[d] + bE * sv.sx[(size_t)k * D + d]);
    mu_n[d] = e1 / e2;
  }
  __syncthreads();
  for (int idx = tid; idx < D * D; idx += nth) {
    const int d1 = idx / D, d2 = idx - d1 * D;
    const double e3 = (1.0 - lrate) * (p[D + idx] + ka_o * mu_o[d1] * mu_o[d2])
                    + lrate * (pr[D + idx] + ka_p * mu_p[d1] * mu_p[d2]
                               + bE * sv.sxx[(size_t)k * D * D + idx]);
    p[D + idx] = e3 - mu_n[d1] * mu_n[d2] * e2;
  }
  __syncthreads();
  for (int d = tid; d < D; d += nth) p[d] = mu_n[d];
  if (tid == 0) { p[D + D * D] = e2; p[D + D * D + 1] = e4 - 2.0 - D; }
}

// Diagonal: one thread per (state, dim); util.py:28-60 with p = 1 per dimension.
__global__ void k_update_emit_svi_diag(int K, int D, double* __restrict__ emit,
                                       const double* __restrict__ prior, const double* __restrict__ stats,
                                       double lrate, double bE) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= K * D) return;
  const int k = e / D, d = e - k * D;
  const StatsView sv = stats_view(stats, K, D, D);
  double* p = emit + (size_t)k * 4 * D;
  const double* pr = prior + (size_t)k * 4 * D;
  const double mu = p[d], sg = p[D + d], ka = p[2 * D + d], nu = p[3 * D + d];
  const double mu0 = pr[d], sg0 = pr[D + d], ka0 = pr[2 * D + d], nu0 = pr[3 * D + d];
  const double nk = sv.n[k];
  const double e1 = (1.0 - lrate) * ka * mu + lrate * (ka0 * mu0 + bE * sv.sx[e]);
  const double e2 = (1.0 - lrate) * ka + lrate * (ka0 + bE * nk);
  const double e3 = (1.0 - lrate) * (sg + ka * mu * mu) + lrate * (sg0 + ka0 * mu0 * mu0 + bE * sv.sxx[e]);
  const double e4 = (1.0 - lrate) * (nu + 3.0) + lrate * (nu0 + 3.0 + bE * nk);
  const double mn = e1 / e2;
  p[d] = mn; p[D + d] = e3 - mn * mn * e2; p[2 * D + d] = e2; p[3 * D + d] = e4 - 3.0;
}

// Batch CAVI: var_init = prior_init + q0 ; var_tran = prior_tran + A.
__global__ void k_update_tran_batch(int K, double* __restrict__ W, double* __restrict__ vinit_user,
                                    const double* __restrict__ prior_tran,
                                    const double* __restrict__ prior_init,
                                    const double* __restrict__ stats, int D, int DD) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const StatsView sv = stats_view(stats, K, D, DD);
  if (idx < K * K) W[idx] = prior_tran[idx] + sv.A[idx];
  if (idx < K) vinit_user[idx] = prior_init[idx] + sv.q0[idx];
}

// Conjugate NIW update from raw moments (n, sx, sxx): xbar = sx/n, centred scatter = sxx - n xbar xbar^T.
// One CTA per state; dynamic smem D doubles.
__global__ void k_update_emit_batch_full(int K, int D, size_t plen, double* __restrict__ emit,
                                         const double* __restrict__ prior,
                                         const double* __restrict__ stats) {
  extern __shared__ double sm[];
  double* xbar = sm;
  const int k = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  const StatsView sv = stats_view(stats, K, D, D * D);
  double* p = emit + (size_t)k * plen;
  const double* pr = prior + (size_t)k * plen;
  const double n = sv.n[k];
  const double ka0 = pr[D + D * D], nu0 = pr[D + D * D + 1];
  if (!(n > SVIHMM_WEPS)) {          // distributions.py:267,275-276: keep the prior
    for (int idx = tid; idx < D + D * D + 2; idx += nth) p[idx] = pr[idx];
    return;
  }
  for (int d = tid; d < D; d += nth) xbar[d] = sv.sx[(size_t)k * D + d] / n;
  __syncthreads();
  for (int idx = tid; idx < D * D; idx += nth) {
    const int d1 = idx / D, d2 = idx - d1 * D;
    const double sumsq = sv.sxx[(size_t)k * D * D + idx] - n * xbar[d1] * xbar[d2];
    p[D + idx] = pr[D + idx] + sumsq + ka0 * n / (ka0 + n) * (xbar[d1] - pr[d1]) * (xbar[d2] - pr[d2]);
  }
  for (int d = tid; d < D; d += nth) p[d] = ka0 / (ka0 + n) * pr[d] + n / (ka0 + n) * xbar[d];
  if (tid == 0) { p[D + D * D] = ka0 + n; p[D + D * D + 1] = nu0 + n; }
}

__global__ void k_update_emit_batch_diag(int K, int D, double* __restrict__ emit,
                                         const double* __restrict__ prior,
                                         const double* __restrict__ stats) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= K * D) return;
  const int k = e / D, d = e - k * D;
  const StatsView sv = stats_view(stats, K, D, D);
  double* p = emit + (size_t)k * 4 * D;
  const double* pr = prior + (size_t)k * 4 * D;
  const double n = sv.n[k];
  const double mu0 = pr[d], sg0 = pr[D + d], ka0 = pr[2 * D + d], nu0 = pr[3 * D + d];
  if (!(n > SVIHMM_WEPS)) { p[d] = mu0; p[D + d] = sg0; p[2 * D + d] = ka0; p[3 * D + d] = nu0; return; }
  const double xb = sv.sx[e] / n;
  const double sumsq = sv.sxx[e] - n * xb * xb;
  p[d] = ka0 / (ka0 + n) * mu0 + n / (ka0 + n) * xb;
  p[D + d] = sg0 + sumsq + ka0 * n / (ka0 + n) * (xb - mu0) * (xb - mu0);
  p[2 * D + d] = ka0 + n;
  p[3 * D + d] = nu0 + n;
}
