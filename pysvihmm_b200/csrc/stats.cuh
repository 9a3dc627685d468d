// K4: expected sufficient statistics as ONE contraction over all rows r = (b,t) of the minibatch
//     Out[m][n] = sum_r left[r][m] * F[r][n],   m < K,  n < N = K + 1 + D + DD
// with the feature row generated on the fly
//     F[r] = [ next[(b,t+1)][0..K) | w | w*x | w*x(x)x ],   w = 1 - mask[r]
// so that columns [0,K) give the transition statistic sum_t outer(q[t], q[t+1]) (with the
// reference's wrap-around pair (T-1,0) if requested: hmmsgd_metaobs.py:876-878, quirks Q1/Q2;
// hmmbatchcd.py:182-184 without) and columns [K,N) the weighted NIW statistics of
// util.NIW_suffstats (util.py:73-83) with masked rows dropped (hmmsgd_metaobs.py:884,903).
// The row range is split over blockIdx.z; fp32 partials are summed in fp64 by k_stats_finalize
// (deterministic, no atomics), which is also where the minibatch accumulation
// hmmsgd_metaobs.py:430-433 happens.
#pragma once
#include "common.cuh"

#define ST_TN 64
#define ST_RC 32

struct StatsArgs {
  int B, T, K, D, DD, N, n_lo, n_hi, diag, wrap, cat;
  int64_t R, rows_per_split;
  const float* left; const float* next;
  const void* obs; int dtype; const uint8_t* mask; const int64_t* starts;
  float* part;
};

// 256 threads = 16 (ty: rows of the tile) x 16 (tx: 4 columns each); dynamic smem ST_RC*D floats
template <int TM>
__global__ void __launch_bounds__(256) k_stats(const StatsArgs a) {
  constexpr int MT = TM / 16;
  __shared__ float Ls[ST_RC][TM];
  __shared__ __align__(16) float Fs[ST_RC][ST_TN];
  __shared__ float wv[ST_RC];
  extern __shared__ float xs[];   // [ST_RC][D]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int K = a.K, D = a.D, T = a.T;
  const int n0 = a.n_lo + blockIdx.x * ST_TN, m0 = blockIdx.y * TM;
  const int64_t rbeg = (int64_t)blockIdx.z * a.rows_per_split;
  const int64_t rend = min(a.R, rbeg + a.rows_per_split);
  const bool need_x = (n0 + ST_TN > K + 1) && (a.n_hi > K + 1);
  float acc[MT][4];
#pragma unroll
  for (int mi = 0; mi < MT; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) acc[mi][ni] = 0.f;

  for (int64_t rc = rbeg; rc < rend; rc += ST_RC) {
    for (int idx = tid; idx < ST_RC * TM; idx += 256) {
      const int rr = idx / TM, m = idx - rr * TM;
      const int64_t r = rc + rr;
      Ls[rr][m] = (r < rend && m0 + m < K) ? a.left[r * K + m0 + m] : 0.f;
    }
    if (need_x && a.cat) {                     // one symbol per row
      for (int rr = tid; rr < ST_RC; rr += 256) {
        const int64_t r = rc + rr;
        float v = 0.f;
        if (r < rend) {
          const int b = (int)(r / T); const int t = (int)(r - (int64_t)b * T);
          v = (float)ld_obs(a.obs, a.dtype, a.starts[b] + t);
        }
        xs[rr] = v;
      }
    } else if (need_x) {
      for (int idx = tid; idx < ST_RC * D; idx += 256) {
        const int rr = idx / D, d = idx - rr * D;
        const int64_t r = rc + rr;
        float v = 0.f;
        if (r < rend) {
          const int b = (int)(r / T); const int t = (int)(r - (int64_t)b * T);
          v = (float)ld_obs(a.obs, a.dtype, (a.starts[b] + t) * D + d);
        }
        xs[rr * D + d] = v;
      }
    }
    __syncthreads();
    if (tid < ST_RC) {
      const int64_t r = rc + tid;
      float w = 0.f;
      if (r < rend) {
        const int b = (int)(r / T); const int t = (int)(r - (int64_t)b * T);
        w = (a.mask && a.mask[a.starts[b] + t]) ? 0.f : 1.f;
        if (need_x && a.cat) {
          const float x = xs[tid];
          if (isnan(x) || x < 0.f || x >= (float)D) { w = 0.f; xs[tid] = -1.f; }
        } else if (need_x) {
          bool bad = false;
          for (int d = 0; d < D; ++d) bad |= isnan(xs[tid * D + d]);
          if (bad || w == 0.f) { w = 0.f; for (int d = 0; d < D; ++d) xs[tid * D + d] = 0.f; }
        }
      }
      wv[tid] = w;
    }
    __syncthreads();
    for (int idx = tid; idx < ST_RC * ST_TN; idx += 256) {
      const int rr = idx / ST_TN, nn = idx - rr * ST_TN;
      const int n = n0 + nn;
      const int64_t r = rc + rr;
      float val = 0.f;
      if (r < rend && n < a.n_hi) {
        if (n < K) {
          const int b = (int)(r / T); int t = (int)(r - (int64_t)b * T) + 1;
          const bool valid = (t < T) || a.wrap;
          if (t == T) t = 0;
          if (valid) val = a.next[((int64_t)b * T + t) * K + n];
        } else if (n == K) {
          val = wv[rr];
        } else if (a.cat) {
          val = (wv[rr] != 0.f && (int)xs[rr] == n - K - 1) ? 1.f : 0.f;     // w * 1[x = c]
        } else if (n < K + 1 + D) {
          val = wv[rr] * xs[rr * D + (n - K - 1)];
        } else {
          const int m = n - K - 1 - D;
          if (a.diag) { const float x = xs[rr * D + m]; val = wv[rr] * x * x; }
          else { const int d1 = m / D, d2 = m - d1 * D; val = wv[rr] * xs[rr * D + d1] * xs[rr * D + d2]; }
        }
      }
      Fs[rr][nn] = val;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < ST_RC; ++rr) {
      const float4 f = *reinterpret_cast<const float4*>(&Fs[rr][tx * 4]);
#pragma unroll
      for (int mi = 0; mi < MT; ++mi) {
        const float l = Ls[rr][ty * MT + mi];
        acc[mi][0] = fmaf(l, f.x, acc[mi][0]);
        acc[mi][1] = fmaf(l, f.y, acc[mi][1]);
        acc[mi][2] = fmaf(l, f.z, acc[mi][2]);
        acc[mi][3] = fmaf(l, f.w, acc[mi][3]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int mi = 0; mi < MT; ++mi)
#pragma unroll
    for (int ni = 0; ni < 4; ++ni) {
      const int m = m0 + ty * MT + mi, n = n0 + tx * 4 + ni;
      if (m < K && n < a.n_hi) a.part[((size_t)blockIdx.z * K + m) * a.N + n] = acc[mi][ni];
    }
}

// Sum the row-split partials in float64 and lay the statistics out as include/svihmm.h documents.
__global__ void __launch_bounds__(256)
k_stats_finalize(int B, int T, int K, int D, int DD, int N, int nsplit, const float* __restrict__ part,
                 const float* __restrict__ q, const double* __restrict__ seq,
                 const double* __restrict__ prior_tran, int add_prior, const float* __restrict__ Pt,
                 int exact_xi, double* __restrict__ out, size_t slen) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= slen) return;
  const size_t o_n = (size_t)K * K, o_sx = o_n + K, o_sxx = o_sx + (size_t)K * D,
               o_q0 = o_sxx + (size_t)K * DD, o_tail = o_q0 + K;
  int m = -1, n = 0;
  double v = 0.0;
  if (idx < o_n) { m = (int)(idx / K); n = (int)(idx % K); }
  else if (idx < o_sx) { m = (int)(idx - o_n); n = K; }
  else if (idx < o_sxx) { const size_t e = idx - o_sx; m = (int)(e / D); n = K + 1 + (int)(e % D); }
  else if (idx < o_q0) { const size_t e = idx - o_sxx; m = (int)(e / DD); n = K + 1 + D + (int)(e % DD); }
  if (m >= 0) {
    for (int z = 0; z < nsplit; ++z) v += (double)part[((size_t)z * K + m) * N + n];
    if (idx < o_n) {
      if (exact_xi) v *= (double)Pt[idx];
      if (add_prior) v += (double)B * (prior_tran[idx] - 1.0);
    }
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

// Mixture layout: transition partials partT [nsplitT][K][K] and component partials
// partE [nsplitE][KE][NE] (NE = KE + 1 + D + DD, columns < KE unused) ->
// [ A (K*K) | n (KE) | sx (KE*D) | sxx (KE*DD) | q0 (K) | tail ].
__global__ void __launch_bounds__(256)
k_stats_finalize_mix(int B, int T, int K, int KE, int D, int DD, int nsplitT, int nsplitE,
                     const float* __restrict__ partT, const float* __restrict__ partE,
                     const float* __restrict__ q, const double* __restrict__ seq,
                     const double* __restrict__ prior_tran, int add_prior, double* __restrict__ out, size_t slen) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= slen) return;
  const int NE = KE + 1 + D + DD;
  const size_t o_n = (size_t)K * K, o_sx = o_n + KE, o_sxx = o_sx + (size_t)KE * D,
               o_q0 = o_sxx + (size_t)KE * DD, o_tail = o_q0 + K;
  double v = 0.0;
  if (idx < o_n) {
    for (int z = 0; z < nsplitT; ++z) v += (double)partT[(size_t)z * K * K + idx];
    if (add_prior) v += (double)B * (prior_tran[idx] - 1.0);
  } else if (idx < o_q0) {
    int m, n;
    if (idx < o_sx) { m = (int)(idx - o_n); n = KE; }
    else if (idx < o_sxx) { const size_t e = idx - o_sx; m = (int)(e / D); n = KE + 1 + (int)(e % D); }
    else { const size_t e = idx - o_sxx; m = (int)(e / DD); n = KE + 1 + D + (int)(e % DD); }
    for (int z = 0; z < nsplitE; ++z) v += (double)partE[((size_t)z * KE + m) * NE + n];
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
