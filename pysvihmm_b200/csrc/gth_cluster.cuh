// Stationary vector of the row-stochastic G = W / rowsum for 64 < K <= 384 (BASELINE config 4) on a
// thread-block CLUSTER with the matrix in distributed shared memory: the reference takes the leading
// eigenvector with numpy's eig per meta-observation (hmmsgd_metaobs.py:413-418; 192 ms at K = 256, SURVEY
// section 8 row a2); gth_block (global.cuh) ran the Grassmann-Taksar-Heyman elimination on ONE CTA with the
// matrix in L2 - K - 1 censoring steps, each waiting for L2 round trips of the rows, and K^3 / 3 x 16 bytes
// through one SM's L2 port: 1.6 ms of the 7.5 ms step at K = 256.
//
//   8 CTAs (one cluster), CTA r owns rows i = r, r + 8, ... in its shared memory (64 KB at K = 256).
//   step n = K-1 .. 1: the owner of row n forms s = sum_{j<n} G[n][j] (one warp, fixed order: every rank of
//   a multi-GPU job gets the same bits) while its other warps push the pivot row into the pivot buffer of all
//   8 CTAs through DSMEM (double buffered by step parity); barrier.cluster; every CTA updates its rows i < n:
//   f = G[i][n] / s (kept in place for the back-substitution), G[i][j] += f G[n][j], j < n.  No subtractions,
//   so the result is componentwise accurate as before.
//   back-substitution: pi[0] = 1, pi[j] += pi[i] G[i][j] row by row on warp 0 of CTA 0, the rows read through
//   DSMEM four ahead; then the L2 normalisation, |.| and the digamma transform (pi0_section), as in block 0 of
//   k_global_step, which skips all of this when a.gth_ext is set - and leaves the digamma transforms of the
//   transition matrix (Pt, PtT) to the 4096 threads of this kernel.
#pragma once
#include <cooperative_groups.h>
#include "global.cuh"

#define GC_CTAS 8
#define GC_NT 512
#define GC_KMAX 384

__host__ __device__ inline size_t gth_cluster_smem(int K) {
  const size_t nr = (size_t)(K + GC_CTAS - 1) / GC_CTAS;
  return (nr * K + 2 * (size_t)(K + 8) + (size_t)K + 8) * sizeof(double);
}

__global__ void __cluster_dims__(GC_CTAS, 1, 1) __launch_bounds__(GC_NT, 1)
k_gth_cluster(const int K, const double* __restrict__ G, const double* __restrict__ W, const double* __restrict__ rowsum,
              float* __restrict__ Pt, float* __restrict__ PtT, double* __restrict__ vinit, float* __restrict__ pi0) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) double gcs[];
  const int nr = (K + GC_CTAS - 1) / GC_CTAS;
  double* rows = gcs;                                   // [nr][K]: local row l = global row l * 8 + rank
  double* pivot = rows + (size_t)nr * K;                // [2][K + 8]: pivot row (j < n), [K] = 1 / s
  double* pi = pivot + 2 * (size_t)(K + 8);             // CTA 0: the vector
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5, nw = GC_NT / 32;
  for (int l = wp; l < nr; l += nw) {
    const int i = l * GC_CTAS + rank;
    for (int j = lane; j < K; j += 32) rows[(size_t)l * K + j] = i < K ? G[(size_t)i * K + j] : 0.0;
  }
  // P = exp(psi(W) - psi(rowsum)) (hmmsgd_metaobs.py:503-504) and its transpose: K^2 digamma pairs over the 4096
  // threads of the cluster instead of the 512 of block 0 of k_global_step (0.24 ms there at K = 256)
  {
    const int KK = K * K;
#pragma unroll 1
    for (int idx = rank * GC_NT + tid; idx < KK; idx += GC_CTAS * GC_NT) {
      const int i = idx / K, j = idx - i * K;
      const float v = (float)dexp_ni(digamma_fast(W[idx] + SVIHMM_EPS) - digamma_fast(rowsum[i] + SVIHMM_EPS));
      Pt[idx] = v;
      PtT[j * K + i] = v;
    }
  }
  cluster.sync();
#pragma unroll 1
  for (int n = K - 1; n >= 1; --n) {
    const int owner = n % GC_CTAS, lo = n / GC_CTAS;
    double* buf = pivot + (size_t)(n & 1) * (K + 8);
    if (rank == owner) {
      const double* rown = rows + (size_t)lo * K;
      if (wp == 0) {
        double s = 0.0;
#pragma unroll 1
        for (int j = lane; j < n; j += 32) s += rown[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        const double rinv = 1.0 / s;
        if (lane < GC_CTAS) cluster.map_shared_rank(buf, lane)[K] = rinv;
      } else {
        // (target CTA, j) pairs over the other 15 warps
        for (int idx = tid - 32; idx < GC_CTAS * n; idx += GC_NT - 32) {
          const int t = idx / n, j = idx - t * n;
          cluster.map_shared_rank(buf, t)[j] = rown[j];
        }
      }
    }
    cluster.sync();                                     // pivot row and 1 / s have landed everywhere
    const double rinv = buf[K];
    for (int l = wp; l < nr; l += nw) {
      const int i = l * GC_CTAS + rank;
      if (i >= n) break;
      double* rowi = rows + (size_t)l * K;
      const double f = rowi[n] * rinv;
#pragma unroll 4
      for (int j = lane; j < n; j += 32) rowi[j] = fma(f, buf[j], rowi[j]);
      __syncwarp();
      if (lane == 0) rowi[n] = f;
    }
    __syncthreads();                                    // row n - 1 of its owner is final before it is summed / sent
  }
  cluster.sync();
  if (rank == 0 && wp == 0) {
    for (int j = lane; j < K; j += 32) pi[j] = j == 0 ? 1.0 : 0.0;
    __syncwarp();
    // row i lives in CTA i % 8, local row i / 8; four rows in flight
    constexpr int PF = 4, NU = GC_KMAX / 32;
    double v[PF][NU];
    auto fetch = [&](const int i, double (&dst)[NU]) {
      const double* src = cluster.map_shared_rank(rows + (size_t)(i / GC_CTAS) * K, i % GC_CTAS);
#pragma unroll
      for (int u = 0; u < NU; ++u) { const int j = lane + 32 * u; dst[u] = (j > i && j < K) ? src[j] : 0.0; }
    };
#pragma unroll
    for (int p = 0; p < PF; ++p) if (p < K - 1) fetch(p, v[p]);
#pragma unroll 1
    for (int i0 = 0; i0 < K - 1; i0 += PF) {
#pragma unroll
      for (int p = 0; p < PF; ++p) {
        const int i = i0 + p;
        if (i < K - 1) {
          const double pv = pi[i];
#pragma unroll
          for (int u = 0; u < NU; ++u) { const int j = lane + 32 * u; if (j > i && j < K) pi[j] = fma(pv, v[p][u], pi[j]); }
          __syncwarp();
          if (i + PF < K - 1) fetch(i + PF, v[p]);
        }
      }
    }
    pi0_section(0, K, pi, vinit, pi0, lane, nullptr);
  }
  cluster.sync();                                       // nobody leaves while CTA 0 still reads remote rows
}
