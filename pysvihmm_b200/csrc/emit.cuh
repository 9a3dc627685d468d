// K1: per-timestep expected log-likelihoods of the observation models, float64 accumulation
// (fp32 rounding of the quadratic form is the precision bottleneck of the whole E-step, SURVEY
// section 7), replacing the per-state loop hmmsgd_metaobs.py:508-509 over
// Gaussian.expected_log_likelihood (pybasicbayes/distributions.py:351-359) + np.nan_to_num.
#pragma once
#include "common.cuh"

#define EMIT_ROWS 128   // rows (b,t) per CTA, one per thread

// ll[r][k] = ck - || Rs_k x_r - gk ||^2 ; x staged in smem as doubles, Rs_k staged per state.
// dynamic smem: (EMIT_ROWS*D + D(D+1)/2 + D) doubles.
__global__ void __launch_bounds__(EMIT_ROWS)
k_emit_full(int B, int T, int K, int D, const void* __restrict__ obs, int dtype,
            const uint8_t* __restrict__ mask, const int64_t* __restrict__ starts, int mask_ll,
            const double* __restrict__ Rs, const double* __restrict__ gk,
            const double* __restrict__ ck, double* __restrict__ ll) {
  extern __shared__ double sm[];
  double* xs = sm;                               // [D][EMIT_ROWS]
  double* Rk = sm + (size_t)EMIT_ROWS * D;       // packed lower
  const int tri = D * (D + 1) / 2;
  double* gg = Rk + tri;                         // [D]
  const int tid = threadIdx.x;
  const int64_t R = (int64_t)B * T;
  const int64_t r0 = (int64_t)blockIdx.x * EMIT_ROWS;
  // stage x (coalesced over the D contiguous values of consecutive rows of one window)
  bool dead = false;     // row carries no evidence: NaN anywhere or masked with MASK_LL
  for (int idx = tid; idx < EMIT_ROWS * D; idx += EMIT_ROWS) {
    const int rr = idx / D, d = idx - rr * D;
    const int64_t r = r0 + rr;
    double v = 0.0;
    if (r < R) {
      const int b = (int)(r / T); const int t = (int)(r - (int64_t)b * T);
      v = ld_obs(obs, dtype, (starts[b] + t) * D + d);
    }
    xs[d * EMIT_ROWS + rr] = v;
  }
  __syncthreads();
  const int64_t r = r0 + tid;
  if (r < R) {
    const int b = (int)(r / T); const int t = (int)(r - (int64_t)b * T);
    if (mask_ll && mask && mask[starts[b] + t]) dead = true;
    for (int d = 0; d < D; ++d) if (isnan(xs[d * EMIT_ROWS + tid])) dead = true;
  }
  for (int k = 0; k < K; ++k) {
    __syncthreads();
    for (int idx = tid; idx < tri; idx += EMIT_ROWS) Rk[idx] = Rs[(size_t)k * tri + idx];
    for (int idx = tid; idx < D; idx += EMIT_ROWS) gg[idx] = gk[(size_t)k * D + idx];
    __syncthreads();
    if (r < R) {
      double acc = 0.0;
      int o = 0;
      for (int i = 0; i < D; ++i) {
        double s = -gg[i];
        for (int j = 0; j <= i; ++j) s = fma(Rk[o + j], xs[j * EMIT_ROWS + tid], s);
        o += i + 1;
        acc = fma(s, s, acc);
      }
      ll[r * K + k] = dead ? 0.0 : ck[k] - acc;
    }
  }
}

// Diagonal emissions: one thread per (row, state); parameters in smem.
// dynamic smem: 2*K*D doubles.
__global__ void __launch_bounds__(256)
k_emit_diag(int B, int T, int K, int D, const void* __restrict__ obs, int dtype,
            const uint8_t* __restrict__ mask, const int64_t* __restrict__ starts, int mask_ll,
            const double* __restrict__ Rs, const double* __restrict__ gk,
            const double* __restrict__ ck, double* __restrict__ ll, int p_smem) {
  extern __shared__ double sm[];
  // parameters in shared memory when 2*K*D doubles fit, else read through L1/L2 (K*D = 16384 at config 4)
  const double* rr_ = p_smem ? sm : Rs; const double* mm_ = p_smem ? sm + (size_t)K * D : gk;
  if (p_smem) {
    for (int idx = threadIdx.x; idx < K * D; idx += blockDim.x) { sm[idx] = Rs[idx]; sm[(size_t)K * D + idx] = gk[idx]; }
    __syncthreads();
  }
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t R = (int64_t)B * T;
  if (e >= R * K) return;
  const int64_t r = e / K; const int k = (int)(e - r * K);
  const int b = (int)(r / T); const int t = (int)(r - (int64_t)b * T);
  const int64_t gi = starts[b] + t;
  bool dead = mask_ll && mask && mask[gi];
  double acc = 0.0;
  for (int d = 0; d < D; ++d) {
    const double x = ld_obs(obs, dtype, gi * D + d);
    if (isnan(x)) dead = true;
    const double df = x - mm_[k * D + d];
    acc = fma(rr_[k * D + d] * df, df, acc);
  }
  ll[e] = dead ? 0.0 : ck[k] - acc;
}

// Diagonal emissions for large K*D (parameters do not fit in shared memory at once): the CTA owns a
// tile of 32 states (their parameters in shared memory: 2*32*D doubles) and a block of rows; a warp =
// the 32 states of ONE row, so the observation loads are warp-uniform broadcasts, and every thread
// keeps ED_R rows in flight so that each parameter load feeds ED_R*2 DFMA.
#define ED_R 4
__global__ void __launch_bounds__(256)
k_emit_diag_tiled(int B, int T, int K, int D, const void* __restrict__ obs, int dtype,
                  const uint8_t* __restrict__ mask, const int64_t* __restrict__ starts, int mask_ll,
                  const double* __restrict__ Rs, const double* __restrict__ gk,
                  const double* __restrict__ ck, double* __restrict__ ll) {
  extern __shared__ double sm[];                  // [D][32] rs, [D][32] mu
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int k0 = blockIdx.y * 32, k = k0 + lane;
  for (int idx = threadIdx.x; idx < 32 * D; idx += blockDim.x) {
    const int d = idx >> 5, kk = idx & 31;
    sm[idx] = k0 + kk < K ? Rs[(size_t)(k0 + kk) * D + d] : 0.0;
    sm[32 * D + idx] = k0 + kk < K ? gk[(size_t)(k0 + kk) * D + d] : 0.0;
  }
  __syncthreads();
  const int64_t R = (int64_t)B * T;
  const double c = k < K ? ck[k] : 0.0;
  for (int64_t r0 = ((int64_t)blockIdx.x * 8 + wp) * ED_R; r0 < R; r0 += (int64_t)gridDim.x * 8 * ED_R) {
    int64_t gi[ED_R]; bool dead[ED_R]; double acc[ED_R];
#pragma unroll
    for (int u = 0; u < ED_R; ++u) {
      const int64_t r = r0 + u;
      gi[u] = 0; dead[u] = r >= R; acc[u] = 0.0;
      if (r < R) {
        const int b = (int)(r / T); const int t = (int)(r - (int64_t)b * T);
        gi[u] = starts[b] + t;
        dead[u] = mask_ll && mask && mask[gi[u]];
      }
    }
    for (int d = 0; d < D; ++d) {
      const double rs = sm[d * 32 + lane], mu = sm[32 * D + d * 32 + lane];
#pragma unroll
      for (int u = 0; u < ED_R; ++u) {
        const double x = ld_obs(obs, dtype, gi[u] * D + d);       // same address on every lane: broadcast
        if (isnan(x)) dead[u] = true;
        const double df = x - mu;
        acc[u] = fma(rs * df, df, acc[u]);
      }
    }
    if (k < K) {
#pragma unroll
      for (int u = 0; u < ED_R; ++u) if (r0 + u < R) ll[(r0 + u) * K + k] = dead[u] ? 0.0 : c - acc[u];
    }
  }
}

// Categorical emissions (pybasicbayes/distributions.py:1383-1386): ll[r][k] = logp[k][x_r] with the
// table logp[k][c] = psi(alpha_mf[k][c]) - psi(sum_c alpha_mf[k][c]) prepared by the global step.
// One thread per (row, state).  A NaN, masked (with mask_ll) or out-of-range symbol gives ll = 0.
__global__ void __launch_bounds__(256)
k_emit_cat(int B, int T, int K, int C, const void* __restrict__ obs, int dtype,
           const uint8_t* __restrict__ mask, const int64_t* __restrict__ starts, int mask_ll,
           const double* __restrict__ logp, double* __restrict__ ll) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t R = (int64_t)B * T;
  if (e >= R * K) return;
  const int64_t r = e / K; const int k = (int)(e - r * K);
  const int b = (int)(r / T); const int t = (int)(r - (int64_t)b * T);
  const int64_t gi = starts[b] + t;
  const double x = ld_obs(obs, dtype, gi);
  const bool dead = (mask_ll && mask && mask[gi]) || isnan(x) || x < 0.0 || x >= (double)C;
  ll[e] = dead ? 0.0 : logp[(size_t)k * C + (int)x];
}

// Mixture emissions (EXTENSION, BASELINE config 5): state k emits from C components with expected
// log-weights lw[k*C + c] = E[ln pi_kc].  From the component expected log-likelihoods ell[r][k*C+c]
// (distributions.py:351-366, computed by the kernels above with K*C "states"):
//     ll[r][k]    = logsumexp_c( lw[kc] + ell[r][kc] )          (labels.py:52-65: logr, then softmax)
//     resp[r][kc] = exp( lw[kc] + ell[r][kc] - ll[r][k] )
// A row without evidence (NaN, or masked with mask_ll) has ell = 0 for every component and gets
// ll = 0 (np.nan_to_num, hmmsgd_metaobs.py:508-509).  One thread per (row, state).
__global__ void __launch_bounds__(256)
k_mix_combine(int B, int T, int K, int C, int D, const void* __restrict__ obs, int dtype,
              const uint8_t* __restrict__ mask, const int64_t* __restrict__ starts, int mask_ll,
              const double* __restrict__ lw, const double* __restrict__ ell, double* __restrict__ ll,
              float* __restrict__ resp) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t R = (int64_t)B * T;
  if (e >= R * K) return;
  const int64_t r = e / K; const int k = (int)(e - r * K);
  const int b = (int)(r / T); const int t = (int)(r - (int64_t)b * T);
  const int64_t gi = starts[b] + t;
  bool dead = mask_ll && mask && mask[gi];
  for (int d = 0; d < D; ++d) dead |= isnan(ld_obs(obs, dtype, gi * D + d));
  const double* ep = ell + (r * K + k) * C;
  const double* lp = lw + (size_t)k * C;
  float* rp = resp + (r * K + k) * C;
  if (dead) {
    ll[e] = 0.0;
    for (int c = 0; c < C; ++c) rp[c] = 0.f;
    return;
  }
  // The differences to the maximum are formed in float64; their exponentials and the logarithm of the sum
  // (in [1, C]) in float32: |error of ll| < 3e-7, far below what moves a marginal by 1e-5, and the float64
  // exp / log (150 / 290 cycles each, 2 C + 1 per thread) were most of this kernel.
  double m = -INFINITY;
  for (int c = 0; c < C; ++c) m = fmax(m, lp[c] + ep[c]);
  float s = 0.f;
  for (int c = 0; c < C; ++c) s += __expf((float)(lp[c] + ep[c] - m));
  const float ls = __logf(s), inv = 1.f / s;
  ll[e] = m + (double)ls;
  for (int c = 0; c < C; ++c) rp[c] = __expf((float)(lp[c] + ep[c] - m)) * inv;
}

// statistics weights of the components: wq[r][kc] = q[r][k] * resp[r][kc]
__global__ void __launch_bounds__(256)
k_mix_weights(int64_t n, int C, const float* __restrict__ q, const float* __restrict__ resp,
              float* __restrict__ wq) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) wq[e] = q[e / C] * resp[e];
}

// b[r][k] = exp(ll[r][k] - max_k ll[r][k]) (fp32), mx[r] = the max (fp64).
// One warp per row, lanes stride over k.
__global__ void __launch_bounds__(256)
k_ll_to_b(int64_t R, int K, const double* __restrict__ ll, float* __restrict__ b,
          double* __restrict__ mx) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= R) return;
  double m = -INFINITY;
  for (int k = lane; k < K; k += 32) m = fmax(m, ll[r * K + k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  for (int k = lane; k < K; k += 32) b[r * K + k] = (float)exp(ll[r * K + k] - m);
  if (lane == 0) mx[r] = m;
}
