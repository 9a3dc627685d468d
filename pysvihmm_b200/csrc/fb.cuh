// K2/K3 (general path): scaled forward / backward recursions + posterior marginals, batched
// over sequences, with the forward table spilled to HBM.  Replaces forward_msgs
// (hmmsgd_metaobs.py:775-803, hmmbase.py:266-295), backward_msgs (:828-855, hmmbase.py:297-320)
// and the marginal computation (:516-519): the normalised vectors here are softmax_k(lalpha[t])
// and softmax_k(lbeta[t]) of the reference's log-domain tables.
//
// Two mappings:
//   k_forward<KP>/k_backward<KP>   K <= 32: one KP-lane group per sequence (KP = pow2 >= K), the
//                                  transition column/row of each lane lives in registers, the
//                                  K-vector is exchanged with warp shuffles.
//   k_forward_wide/k_backward_wide 32 < K <= ~232: one CTA per sequence, P in shared memory.
#pragma once
#include "common.cuh"

template <int KP>
__device__ __forceinline__ float gsum(float v, unsigned gmask) {
#pragma unroll
  for (int o = KP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o, KP);
  return v;
}

#define FB_U 8   // timesteps whose (independent) loads are issued ahead of the dependent chain

template <int KP>
__global__ void __launch_bounds__(128)
k_forward(int B, int T, int K, const float* __restrict__ Pt, const float* __restrict__ pi0,
          const float* __restrict__ b, float* __restrict__ alpha, float* __restrict__ cs) {
  constexpr int G = 32 / KP;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int grp = lane / KP, j = lane % KP;
  const int s = warp * G + grp;
  if (s >= B) return;
  const unsigned gmask = (KP == 32) ? 0xffffffffu : (((1u << KP) - 1u) << (grp * KP));
  const bool act = j < K;
  float col[KP];
#pragma unroll
  for (int i = 0; i < KP; ++i) col[i] = (act && i < K) ? Pt[i * K + j] : 0.f;
  const size_t base = (size_t)s * T * K + (act ? j : 0);
  const float* bp = b + base;
  float* ap = alpha + base;
  float* cp = cs + (size_t)s * T;
  float a = act ? pi0[j] * bp[0] : 0.f;
  float sum = gsum<KP>(a, gmask);
  a *= 1.f / sum;
  if (act) ap[0] = a;
  if (j == 0) cp[0] = sum;
  for (int t0 = 1; t0 < T; t0 += FB_U) {
    float bb[FB_U];
#pragma unroll
    for (int u = 0; u < FB_U; ++u) bb[u] = (act && t0 + u < T) ? bp[(size_t)(t0 + u) * K] : 0.f;
#pragma unroll
    for (int u = 0; u < FB_U; ++u) {
      const int t = t0 + u;
      if (t < T) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int i = 0; i < KP; i += 4) {
          a0 = fmaf(__shfl_sync(gmask, a, i, KP), col[i], a0);
          if (KP > 1) a1 = fmaf(__shfl_sync(gmask, a, i + 1, KP), col[(i + 1) % KP], a1);
          if (KP > 2) a2 = fmaf(__shfl_sync(gmask, a, i + 2, KP), col[(i + 2) % KP], a2);
          if (KP > 2) a3 = fmaf(__shfl_sync(gmask, a, i + 3, KP), col[(i + 3) % KP], a3);
        }
        const float v = ((a0 + a1) + (a2 + a3)) * bb[u];
        sum = gsum<KP>(v, gmask);
        a = v * (1.f / sum);
        if (act) ap[(size_t)t * K] = a;
        if (j == 0) cp[t] = sum;
      }
    }
  }
}

// beta-hat on the fly; q[t] = normalise(alpha-hat[t] * beta-hat[t]).  With r_out != NULL also
// writes r[t] = b[t]*beta-hat[t] / z[t] (t >= 1) so that the exact pairwise statistic is
// P .* sum_t alpha-hat[t-1] r[t]^T  (SVIHMM_EXACT_XI; not reference behaviour).
template <int KP>
__global__ void __launch_bounds__(128)
k_backward(int B, int T, int K, const float* __restrict__ Pt, const float* __restrict__ b,
           const float* __restrict__ alpha, float* __restrict__ q, float* __restrict__ r_out,
           float* __restrict__ beta_out = nullptr, float* __restrict__ sb_out = nullptr) {
  constexpr int G = 32 / KP;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int grp = lane / KP, i = lane % KP;
  const int s = warp * G + grp;
  if (s >= B) return;
  const unsigned gmask = (KP == 32) ? 0xffffffffu : (((1u << KP) - 1u) << (grp * KP));
  const bool act = i < K;
  float row[KP];
#pragma unroll
  for (int jj = 0; jj < KP; ++jj) row[jj] = (act && jj < K) ? Pt[i * K + jj] : 0.f;
  const size_t base = (size_t)s * T * K + (act ? i : 0);
  const float* bp = b + base;
  const float* ap = alpha + base;
  float* qp = q + base;
  float* rp = r_out ? r_out + base : nullptr;
  float beta = act ? 1.f : 0.f;
  if (act) qp[(size_t)(T - 1) * K] = ap[(size_t)(T - 1) * K];
  // KEEP_LOCALS: the normalised backward messages and their scale factors (self.lbeta, :828-855)
  float* bop = beta_out ? beta_out + base : nullptr;
  float* sbp = sb_out ? sb_out + (size_t)s * T : nullptr;
  if (bop && act) bop[(size_t)(T - 1) * K] = 1.f;
  if (sbp && i == 0) sbp[T - 1] = 1.f;
  for (int t0 = T - 2; t0 >= 0; t0 -= FB_U) {
    float bb[FB_U], aa[FB_U];
#pragma unroll
    for (int u = 0; u < FB_U; ++u) {
      const int t = t0 - u;
      const bool ok = act && t >= 0;
      bb[u] = ok ? bp[(size_t)(t + 1) * K] : 0.f;
      aa[u] = ok ? ap[(size_t)t * K] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < FB_U; ++u) {
      const int t = t0 - u;
      if (t >= 0) {
        const float uu = beta * bb[u];
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int jj = 0; jj < KP; jj += 4) {
          a0 = fmaf(__shfl_sync(gmask, uu, jj, KP), row[jj], a0);
          if (KP > 1) a1 = fmaf(__shfl_sync(gmask, uu, jj + 1, KP), row[(jj + 1) % KP], a1);
          if (KP > 2) a2 = fmaf(__shfl_sync(gmask, uu, jj + 2, KP), row[(jj + 2) % KP], a2);
          if (KP > 2) a3 = fmaf(__shfl_sync(gmask, uu, jj + 3, KP), row[(jj + 3) % KP], a3);
        }
        const float acc = (a0 + a1) + (a2 + a3);
        const float qv = aa[u] * acc;
        const float S1 = gsum<KP>(acc, gmask);
        const float S2 = gsum<KP>(qv, gmask);
        beta = acc * (1.f / S1);
        if (act) {
          qp[(size_t)t * K] = qv * (1.f / S2);
          if (rp) rp[(size_t)(t + 1) * K] = uu * (1.f / S2);
          if (bop) bop[(size_t)t * K] = beta;
        }
        if (sbp && i == 0) sbp[t] = S1;
      }
    }
  }
  if (rp && act) rp[0] = 0.f;
}

// ---- wide variants: one CTA (KT = roundup32(K) threads) per sequence -------------------------
__device__ __forceinline__ void block_sum2(float& v1, float& v2, float* red, int nwarp) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    v1 += __shfl_xor_sync(0xffffffffu, v1, o);
    v2 += __shfl_xor_sync(0xffffffffu, v2, o);
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { red[2 * w] = v1; red[2 * w + 1] = v2; }
  __syncthreads();
  float s1 = 0.f, s2 = 0.f;
  for (int x = 0; x < nwarp; ++x) { s1 += red[2 * x]; s2 += red[2 * x + 1]; }
  v1 = s1; v2 = s2;
}

// dynamic smem: (K*K + 2*KT + 64) floats; with p_smem == 0 (K*K floats do not fit: K > 232) the
// transition matrix is read through L1/L2 instead and the K*K floats are not allocated
__global__ void k_forward_wide(int B, int T, int K, const float* __restrict__ Pt,
                               const float* __restrict__ pi0, const float* __restrict__ b,
                               float* __restrict__ alpha, float* __restrict__ cs, int p_smem) {
  extern __shared__ float smf[];
  const int KT = blockDim.x, nw = KT >> 5, j = threadIdx.x, s = blockIdx.x;
  float* ab = smf + (p_smem ? K * K : 0); float* red = ab + 2 * KT;
  const float* Ps = p_smem ? smf : Pt;
  if (p_smem) for (int idx = j; idx < K * K; idx += KT) smf[idx] = Pt[idx];
  const bool act = j < K;
  const size_t base = (size_t)s * T * K + (act ? j : 0);
  const float* bp = b + base; float* ap = alpha + base;
  float a = act ? pi0[j] * bp[0] : 0.f, dummy = 0.f;
  float sum = a;
  block_sum2(sum, dummy, red, nw);
  a *= 1.f / sum;
  if (act) ap[0] = a;
  ab[j] = a;
  if (j == 0) cs[(size_t)s * T] = sum;
  __syncthreads();
  for (int t = 1; t < T; ++t) {
    const float* av = ab + ((t - 1) & 1) * KT;
    const float bt = act ? bp[(size_t)t * K] : 0.f;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (act) {
      int i = 0;
      for (; i + 3 < K; i += 4) {
        a0 = fmaf(av[i], Ps[i * K + j], a0);
        a1 = fmaf(av[i + 1], Ps[(i + 1) * K + j], a1);
        a2 = fmaf(av[i + 2], Ps[(i + 2) * K + j], a2);
        a3 = fmaf(av[i + 3], Ps[(i + 3) * K + j], a3);
      }
      for (; i < K; ++i) a0 = fmaf(av[i], Ps[i * K + j], a0);
    }
    const float v = ((a0 + a1) + (a2 + a3)) * bt;
    sum = v; dummy = 0.f;
    block_sum2(sum, dummy, red, nw);
    a = v * (1.f / sum);
    if (act) ap[(size_t)t * K] = a;
    ab[(t & 1) * KT + j] = a;
    if (j == 0) cs[(size_t)s * T + t] = sum;
    __syncthreads();
  }
}

// dynamic smem as k_forward_wide; PtT is the transposed transition matrix
__global__ void k_backward_wide(int B, int T, int K, const float* __restrict__ PtT,
                                const float* __restrict__ b, const float* __restrict__ alpha,
                                float* __restrict__ q, float* __restrict__ r_out, int p_smem,
                                float* __restrict__ beta_out = nullptr, float* __restrict__ sb_out = nullptr) {
  extern __shared__ float smf[];
  const int KT = blockDim.x, nw = KT >> 5, i = threadIdx.x, s = blockIdx.x;
  float* ub = smf + (p_smem ? K * K : 0); float* red = ub + 2 * KT;
  const float* PsT = p_smem ? smf : PtT;
  if (p_smem) for (int idx = i; idx < K * K; idx += KT) smf[idx] = PtT[idx];
  const bool act = i < K;
  const size_t base = (size_t)s * T * K + (act ? i : 0);
  const float* bp = b + base; const float* ap = alpha + base;
  float* qp = q + base; float* rp = r_out ? r_out + base : nullptr;
  float beta = act ? 1.f : 0.f;
  if (act) qp[(size_t)(T - 1) * K] = ap[(size_t)(T - 1) * K];
  float* bop = beta_out ? beta_out + base : nullptr;
  float* sbp = sb_out ? sb_out + (size_t)s * T : nullptr;
  if (bop && act) bop[(size_t)(T - 1) * K] = 1.f;
  if (sbp && i == 0) sbp[T - 1] = 1.f;
  for (int t = T - 2; t >= 0; --t) {
    float* uv = ub + (t & 1) * KT;
    const float uu = act ? beta * bp[(size_t)(t + 1) * K] : 0.f;
    const float at = act ? ap[(size_t)t * K] : 0.f;
    uv[i] = uu;
    __syncthreads();
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (act) {
      int jj = 0;
      for (; jj + 3 < K; jj += 4) {
        a0 = fmaf(uv[jj], PsT[jj * K + i], a0);
        a1 = fmaf(uv[jj + 1], PsT[(jj + 1) * K + i], a1);
        a2 = fmaf(uv[jj + 2], PsT[(jj + 2) * K + i], a2);
        a3 = fmaf(uv[jj + 3], PsT[(jj + 3) * K + i], a3);
      }
      for (; jj < K; ++jj) a0 = fmaf(uv[jj], PsT[jj * K + i], a0);
    }
    const float acc = (a0 + a1) + (a2 + a3);
    const float qv = at * acc;
    float S1 = acc, S2 = qv;
    block_sum2(S1, S2, red, nw);     // contains a __syncthreads: red is free again afterwards
    beta = acc * (1.f / S1);
    if (act) {
      qp[(size_t)t * K] = qv * (1.f / S2);
      if (rp) rp[(size_t)(t + 1) * K] = uu * (1.f / S2);
      if (bop) bop[(size_t)t * K] = beta;
    }
    if (sbp && i == 0) sbp[t] = S1;
  }
  if (rp && act) rp[0] = 0.f;
}

// per-sequence log normalisers from the forward scale factors:
//   seq[2s]   = logZ = sum_t (log c_t + mx_t)            ( = logsumexp_k lalpha[T-1,k] )
//   seq[2s+1] = sum_t sum_{u<=t} (log c_u + mx_u)         ( = local_lower_bound, quirk Q4,
//                                                           hmmsgd_metaobs.py:257-271 )
__global__ void __launch_bounds__(256)
k_seq_logz(int B, int T, const float* __restrict__ cs, const double* __restrict__ mx,
           double* __restrict__ seq) {
  const int lane = threadIdx.x & 31;
  const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (s >= B) return;
  double lz = 0.0, q4 = 0.0;
  for (int t = lane; t < T; t += 32) {
    const double term = log((double)cs[(size_t)s * T + t]) + mx[(size_t)s * T + t];
    lz += term;
    q4 += (double)(T - t) * term;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lz += __shfl_xor_sync(0xffffffffu, lz, o);
    q4 += __shfl_xor_sync(0xffffffffu, q4, o);
  }
  if (lane == 0) { seq[2 * s] = lz; seq[2 * s + 1] = q4; }
}

// the same for LONG sequences: grid (B, nsplit), every CTA sums a slice of the rows and adds its two
// partial sums to seq with float64 atomics (seq zeroed beforehand); one warp per sequence would walk
// 1e6 rows alone
__global__ void __launch_bounds__(256)
k_seq_logz_split(int B, int T, const float* __restrict__ cs, const double* __restrict__ mx,
                 double* __restrict__ seq) {
  __shared__ double red[2][8];
  const int s = blockIdx.x, tid = threadIdx.x;
  const int per = (T + gridDim.y - 1) / gridDim.y;
  const int ta = blockIdx.y * per, tb = min(T, ta + per);
  double lz = 0.0, q4 = 0.0;
  for (int t = ta + tid; t < tb; t += 256) {
    const double term = log((double)cs[(size_t)s * T + t]) + mx[(size_t)s * T + t];
    lz += term;
    q4 += (double)(T - t) * term;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lz += __shfl_xor_sync(0xffffffffu, lz, o);
    q4 += __shfl_xor_sync(0xffffffffu, q4, o);
  }
  if ((tid & 31) == 0) { red[0][tid >> 5] = lz; red[1][tid >> 5] = q4; }
  __syncthreads();
  if (tid == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < 8; ++i) { a += red[0][i]; b += red[1][i]; }
    atomicAdd(seq + 2 * s, a); atomicAdd(seq + 2 * s + 1, b);
  }
}
