// Forward-filter backward-SAMPLING of a state path (the reference's only native kernel on the HMM
// side: hmm_fast.pyx:43-124, bound as VariationalHMMBase.ffbs_fast at hmmbase.py:410-411; numpy
// version hmmbase.py:231-264).  Semantics follow hmm_fast.pyx:
//   forward filter   lalpha[0] = mod_init + ll[0],  mod_init = psi(v+eps) - psi(sum v + eps)   (:82-83)
//                    lalpha[t] = logsumexp_i(lalpha[t-1,i] + log(A[i,j] + eps)) + ll[t]          (:97-100)
//                    -- the transition weights are the raw Dirichlet parameters A = var_tran, NOT
//                    exp(E[log A]) (a quirk of the reference kept here)
//   backward sample  z[T-1] ~ softmax(lalpha[T-1]);  z[t] ~ softmax_k(lalpha[t,k] + log(A[k,z[t+1]] + eps))
// The forward filter reuses the scaled forward kernels (fb.cuh) with P' = A + eps and pi0' = exp(mod_init);
// the sampler below draws from alpha~[t,k] * P'[k][z[t+1]] (identical distribution).  libc rand() of
// the reference (:30) is replaced by a counter-based Philox stream per sample, so paths agree with the
// reference in distribution only.
#pragma once
#include <curand_kernel.h>
#include "common.cuh"

// P'[i][j] = A[i][j] + eps (float), pi0'[j] = exp(psi(v_j + eps) - psi(sum v + eps))
__global__ void k_ffbs_prep(int K, const double* __restrict__ W, const double* __restrict__ v,
                            float* __restrict__ Pf, float* __restrict__ pi0f) {
  const double eps = 2.220446049250313e-16;       // DBL_EPSILON, hmm_fast.pyx:14,82-86
  for (int i = threadIdx.x; i < K * K; i += blockDim.x) Pf[i] = (float)(W[i] + eps);
  __shared__ double dgs;
  if (threadIdx.x == 0) { double s = 0.0; for (int j = 0; j < K; ++j) s += v[j]; dgs = digamma_d(s + eps); }
  __syncthreads();
  for (int j = threadIdx.x; j < K; j += blockDim.x) pi0f[j] = (float)exp(digamma_d(v[j] + eps) - dgs);
}

// One warp per sample path.  alpha: (T, K) normalised forward messages of ONE sequence.
__global__ void __launch_bounds__(128)
k_ffbs_sample(int nsamples, int T, int K, const float* __restrict__ alpha, const float* __restrict__ Pf,
              unsigned long long seed, int* __restrict__ z) {
  const int lane = threadIdx.x & 31;
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (s >= nsamples) return;
  curandStatePhilox4_32_10_t rng;
  curand_init(seed, (unsigned long long)s, 0ull, &rng);     // same stream on every lane of the warp
  int znext = -1;
  for (int t = T - 1; t >= 0; --t) {
    const float u = curand_uniform(&rng);                     // (0, 1], identical on all lanes
    // weights w_k = alpha[t,k] * P'[k][znext] in chunks of 32 states; pick the first k whose running sum >= u * total
    float total = 0.f;
    for (int k0 = 0; k0 < K; k0 += 32) {
      const int k = k0 + lane;
      float w = k < K ? alpha[(size_t)t * K + k] * (znext >= 0 ? Pf[(size_t)k * K + znext] : 1.f) : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
      total += w;
    }
    const float target = u * total;
    float run = 0.f;
    int pick = K - 1;
    bool found = false;
    for (int k0 = 0; k0 < K && !found; k0 += 32) {
      const int k = k0 + lane;
      const float w = k < K ? alpha[(size_t)t * K + k] * (znext >= 0 ? Pf[(size_t)k * K + znext] : 1.f) : 0.f;
      float c = w;                                              // inclusive scan over the lanes
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const float n = __shfl_up_sync(0xffffffffu, c, o); if (lane >= o) c += n; }
      c += run;
      const unsigned hit = __ballot_sync(0xffffffffu, k < K && c >= target && w > 0.f);
      if (hit) { pick = k0 + __ffs(hit) - 1; found = true; }
      run = __shfl_sync(0xffffffffu, c, 31);
    }
    znext = pick;
    if (lane == 0) z[(size_t)s * T + t] = pick;
  }
}
