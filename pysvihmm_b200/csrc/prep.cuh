// Globals -> per-step constants, all float64, tiny single-wave kernels.
//   k_prep_tran       hmmsgd_metaobs.py:413-418 (stationary |eigenvector|, quirk Q3),
//                     :502-504 (digamma transforms)
//   k_prep_emit_full  pybasicbayes/distributions.py:351-366 (Cholesky, E[log|Lambda|], constant)
//   k_prep_emit_diag  the same formula per dimension with D = 1 (extension, BASELINE config 2)
#pragma once
#include "common.cuh"

// One CTA.  lu is K*(K+1) doubles of scratch, rowsum K doubles.
__global__ void k_prep_tran(int K, const double* __restrict__ W, const double* __restrict__ user_init,
                            int has_user_init, double* __restrict__ lu, double* __restrict__ rowsum,
                            double* __restrict__ vinit, float* __restrict__ Pt,
                            float* __restrict__ PtT, float* __restrict__ pi0) {
  const int tid = threadIdx.x, nth = blockDim.x;
  __shared__ int s_piv;
  __shared__ double s_red[2];
  for (int i = tid; i < K; i += nth) {
    double s = 0.0;
    for (int j = 0; j < K; ++j) s += W[i * K + j];
    rowsum[i] = s;
  }
  __syncthreads();
  for (int idx = tid; idx < K * K; idx += nth) {
    const int i = idx / K, j = idx - i * K;
    const double v = exp(digamma_d(W[idx] + SVIHMM_EPS) - digamma_d(rowsum[i] + SVIHMM_EPS));
    Pt[idx] = (float)v;
    PtT[j * K + i] = (float)v;
  }
  if (!has_user_init) {
    // Solve (A_mean^T - I) v = 0 with the last equation replaced by sum(v) = 1: the Perron vector
    // np.linalg.eig returns for the eigenvalue 1 (up to scale; rescaled to unit L2 norm below).
    const int ld = K + 1;
    for (int idx = tid; idx < K * K; idx += nth) {
      const int r = idx / K, c = idx - r * K;
      double v = W[c * K + r] / rowsum[c] - (r == c ? 1.0 : 0.0);
      if (r == K - 1) v = 1.0;
      lu[r * ld + c] = v;
    }
    for (int r = tid; r < K; r += nth) lu[r * ld + K] = (r == K - 1) ? 1.0 : 0.0;
    __syncthreads();
    for (int p = 0; p < K; ++p) {
      if (tid == 0) {
        int best = p; double bv = fabs(lu[p * ld + p]);
        for (int r = p + 1; r < K; ++r) { const double a = fabs(lu[r * ld + p]); if (a > bv) { bv = a; best = r; } }
        s_piv = best;
      }
      __syncthreads();
      const int piv = s_piv;
      if (piv != p)
        for (int c = p + tid; c <= K; c += nth) { const double t = lu[p * ld + c]; lu[p * ld + c] = lu[piv * ld + c]; lu[piv * ld + c] = t; }
      __syncthreads();
      const int nr = K - p - 1, nc = K - p;        // rows below the pivot, columns right of it (+rhs)
      const double ppv = lu[p * ld + p];
      for (int idx = tid; idx < nr * nc; idx += nth) {
        const int r = p + 1 + idx / nc, c = p + 1 + idx % nc;
        lu[r * ld + c] -= (lu[r * ld + p] / ppv) * lu[p * ld + c];
      }
      __syncthreads();
    }
    if (tid == 0) {
      double n2 = 0.0, n1 = 0.0;
      for (int r = K - 1; r >= 0; --r) {
        double s = lu[r * ld + K];
        for (int c = r + 1; c < K; ++c) s -= lu[r * ld + c] * vinit[c];
        vinit[r] = s / lu[r * ld + r];
      }
      for (int r = 0; r < K; ++r) { vinit[r] = fabs(vinit[r]); n2 += vinit[r] * vinit[r]; }
      n2 = sqrt(n2);
      for (int r = 0; r < K; ++r) { vinit[r] /= n2; n1 += vinit[r]; }
      s_red[0] = n1;
    }
  } else {
    if (tid == 0) {
      double n1 = 0.0;
      for (int r = 0; r < K; ++r) { vinit[r] = user_init[r]; n1 += user_init[r]; }
      s_red[0] = n1;
    }
  }
  __syncthreads();
  const double dgs = digamma_d(s_red[0] + SVIHMM_EPS);
  for (int i = tid; i < K; i += nth) pi0[i] = (float)exp(digamma_d(vinit[i] + SVIHMM_EPS) - dgs);
}

// One CTA per state; dynamic smem 2*D*D doubles.  Writes
//   Rs[k] (packed lower, D(D+1)/2) = sqrt(nu/2) * chol(sigma)^-1,  gk[k] = Rs m,  ck[k]
// so that  ll = ck - || Rs x - gk ||^2 .
__global__ void k_prep_emit_full(int D, size_t plen, const double* __restrict__ emit,
                                 double* __restrict__ Rs, double* __restrict__ gk,
                                 double* __restrict__ ck) {
  extern __shared__ double sm[];
  double* L = sm;             // D*D
  double* Ri = sm + D * D;    // D*D
  const int k = blockIdx.x, tid = threadIdx.x, nth = blockDim.x;
  const double* p = emit + (size_t)k * plen;
  const double* mu = p; const double* sig = p + D;
  const double kappa = p[D + D * D], nu = p[D + D * D + 1];
  for (int idx = tid; idx < D * D; idx += nth) { L[idx] = sig[idx]; Ri[idx] = 0.0; }
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
  // column c of L^-1 by forward substitution
  for (int c = tid; c < D; c += nth) {
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
    if (j <= i) Rs[k * tri + (size_t)i * (i + 1) / 2 + j] = sc * Ri[idx];
  }
  for (int i = tid; i < D; i += nth) {
    double s = 0.0;
    for (int j = 0; j <= i; ++j) s += sc * Ri[i * D + j] * mu[j];
    gk[(size_t)k * D + i] = s;
  }
  if (tid == 0) {
    double ld = 0.0, dg = 0.0;
    for (int d = 0; d < D; ++d) { ld += log(L[d * D + d]); dg += digamma_d(0.5 * (nu - d)); }
    ck[k] = 0.5 * (dg + D * M_LN2 - 2.0 * ld) - D / (2.0 * kappa) - 0.5 * D * log(2.0 * M_PI);
  }
}

// Diagonal: Rs[k][d] = nu_d / (2 sigma_d), gk[k][d] = mu_d, so ll = ck - sum_d Rs (x_d - mu_d)^2.
__global__ void k_prep_emit_diag(int K, int D, const double* __restrict__ emit, double* __restrict__ Rs,
                                 double* __restrict__ gk, double* __restrict__ ck) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  const double* p = emit + (size_t)k * 4 * D;
  double c = 0.0;
  for (int d = 0; d < D; ++d) {
    const double mu = p[d], sg = p[D + d], ka = p[2 * D + d], nu = p[3 * D + d];
    Rs[(size_t)k * D + d] = nu / (2.0 * sg);
    gk[(size_t)k * D + d] = mu;
    c += 0.5 * (digamma_d(0.5 * nu) + M_LN2 - log(sg)) - 1.0 / (2.0 * ka) - 0.5 * log(2.0 * M_PI);
  }
  ck[k] = c;
}
