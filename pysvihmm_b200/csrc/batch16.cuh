// Batched E-step for K <= 16 (diagonal model): SIXTEEN windows advance together in one warp.
//
// The pipelined single-CTA-per-window kernel (fused_pipe.cuh) runs one dependent chain of T matvecs
// per warp at ~120 cycles per step, so a GPU holds at most ~300 chains and the batched
// forward-backward saturates near 5 % of the HBM roofline however large the minibatch is.  Here the
// per-step matvecs of 16 windows are ONE (16 x 16).(16 x 16) product: rows = windows, contraction =
// source state, columns = destination state, issued as mma.m16n8k8 with every operand split hi/lo
// ("3xTF32": float32-level accuracy, float32 accumulate).  The output fragment of step t is, register
// for register, the A fragment of step t+1 once the contraction index is permuted
// (k-step s, column c  <->  state 8s + 2c [+1]), so the recursion never leaves the registers: no
// shuffles, no shared memory, no per-step normaliser (messages are rescaled by exact powers of two
// chosen from the row maximum seen two steps earlier, as in fused.cuh).  12 MMAs advance 16 chains,
// i.e. 8x the chains per warp-step of the scalar kernels at about the same step latency.
//
// Three launches per minibatch, the (B, T, 16) float32 tables b / alpha / beta pass through L2:
//   k_b16_emit   expected log-likelihoods in float64 (pybasicbayes/distributions.py:351-366 through the
//                per-(d,k) constants of global.cuh), np.nan_to_num semantics (hmmsgd_metaobs.py:508-509),
//                b = exp(ll - max_k ll) float32, row maxima float64; one thread per window row
//   k_b16_chain  forward (hmmsgd_metaobs.py:775-803) and backward (:828-855) recursions, one warp per
//                (group of 16 windows, direction)
//   k_b16_post   marginals (:516-519), per-row log normalisers (:257-271), transition statistic with
//                the reference's wrap-around (:873-881) and the weighted NIW statistics (util.py:73-83)
//                on the tensor cores (3xTF32), summed over the minibatch (:430-433) in float64
// A tcgen05 form (128 windows per CTA as the M dimension) was considered and rejected for this K:
// one step is 48 tensor-pipe cycles but needs a TMEM -> registers -> shared-memory round trip of
// several hundred cycles per step, so it only wins with >= 3 tiles (384 windows) in flight per SM,
// i.e. minibatches of > 50 000 windows.
#pragma once
#include "fused_pipe.cuh"

#define B16_KS 16
#define B16_TGT 167          // target biased exponent of the row maximum: 2^40 (alpha*beta stays < 2^90)
#define B16_PF 8             // chain steps between the load of a b row and its use

struct B16Args {
  int B, T, K, D, wrap, add_prior, mask_ll;
  const void* obs; int dtype; const uint8_t* mask; const int64_t* starts;
  const float *P, *PT, *pi0;
  const double *par2, *ckp, *prior_tran;
  float *bt, *at, *ct; int* Et; double* mx;       // (B*T, 16) tables, (B*T) exponents / row maxima
  float* var_x_out; double* stats_out; double* seq;
  size_t o_n, o_sx, o_sxx, o_q0, o_tail, slen;
};

// ------------------------------------------------------------------------------------------------
// phase A: emissions.  One thread per window row; the 2*D*KP constants of the diagonal model live in
// shared memory as double2 (c2 = -nu/(2 beta'), c1 = 2 c2 mu) and are read as broadcast LDS.128.
// Also zeroes the packed statistics and the per-window log normalisers for the atomics of phase C.
// ------------------------------------------------------------------------------------------------
template <int KP>
__global__ void __launch_bounds__(256) k_b16_emit(const B16Args a) {
  extern __shared__ __align__(16) double b16_par[];          // [D][KP] double2, then ck[KP]
  const int K = a.K, D = a.D, T = a.T;
  const int tid = threadIdx.x;
  for (int i = tid; i < KP * D; i += 256) {
    const int d = i / KP, kk = i - d * KP;
    const bool in = kk < K;
    b16_par[2 * i] = in ? a.par2[2 * (d * K + kk)] : 0.0;
    b16_par[2 * i + 1] = in ? a.par2[2 * (d * K + kk) + 1] : 0.0;
  }
  for (int kk = tid; kk < KP; kk += 256) b16_par[2 * KP * D + kk] = kk < K ? a.ckp[kk] : 0.0;
  const int64_t R = (int64_t)a.B * T;
  const int64_t gtid = (int64_t)blockIdx.x * 256 + tid, nth = (int64_t)gridDim.x * 256;
  for (int64_t i = gtid; i < (int64_t)a.slen; i += nth) a.stats_out[i] = 0.0;
  for (int64_t i = gtid; i < 2 * (int64_t)a.B; i += nth) a.seq[i] = 0.0;
  __syncthreads();
  const bool vecx = a.dtype == SVIHMM_F32 && (D & 3) == 0 && ((((uintptr_t)a.obs) & 15) == 0);
  const bool small = R < (int64_t)0x7fffffff;                 // 32-bit divisions where the row index fits
  for (int64_t r = gtid; r < R; r += nth) {
    const int w = small ? (int)((unsigned)r / (unsigned)T) : (int)(r / T);
    const int t = (int)(r - (int64_t)w * T);
    const int64_t gi = a.starts[w] + t, e0 = gi * D;
    bool bad = false;
    double ll[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) ll[k] = b16_par[2 * KP * D + k];
    for (int d0 = 0; d0 < D; d0 += 4) {
      double xq[4];
      if (vecx) {
        const float4 q = __ldg(reinterpret_cast<const float4*>((const float*)a.obs + e0 + d0));
        xq[0] = q.x; xq[1] = q.y; xq[2] = q.z; xq[3] = q.w;
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) xq[u] = d0 + u < D ? ld_obs(a.obs, a.dtype, e0 + d0 + u) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (d0 + u < D) {
          const double xd = xq[u], xx = xd * xd;
          bad |= isnan(xd);
          const double2* pp = reinterpret_cast<const double2*>(b16_par) + (size_t)(d0 + u) * KP;
#pragma unroll
          for (int k = 0; k < KP; ++k) {
            const double2 c = pp[k];
            ll[k] = fma(c.x, xx, fma(c.y, xd, ll[k]));
          }
        }
      }
    }
    const bool mk = a.mask && a.mask[gi];
    const bool noev = bad || (a.mask_ll && mk);
    double m = -INFINITY;
#pragma unroll
    for (int k = 0; k < KP; ++k) { if (k < K) { if (noev) ll[k] = 0.0; m = fmax(m, ll[k]); } }
    a.mx[r] = m;
    float4* bp = reinterpret_cast<float4*>(a.bt + r * B16_KS);
#pragma unroll
    for (int k4 = 0; k4 < B16_KS; k4 += 4) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k4 < KP) {
        o.x = k4 < K ? __expf((float)(ll[k4] - m)) : 0.f;
        o.y = k4 + 1 < K ? __expf((float)(ll[k4 + 1 < KP ? k4 + 1 : 0] - m)) : 0.f;
        o.z = k4 + 2 < K ? __expf((float)(ll[k4 + 2 < KP ? k4 + 2 : 0] - m)) : 0.f;
        o.w = k4 + 3 < K ? __expf((float)(ll[k4 + 3 < KP ? k4 + 3 : 0] - m)) : 0.f;
      }
      bp[k4 >> 2] = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// phase B: the two recursions.  Fragment conventions of mma.m16n8k8 (g = lane / 4, c = lane % 4):
//   A (16 x 8):  a0 (row g, col c)  a1 (row g+8, col c)  a2 (row g, col c+4)  a3 (row g+8, col c+4)
//   B (8 x 8):   b0 (row c, col g)  b1 (row c+4, col g)
//   D (16 x 8):  d0 (row g, col 2c) d1 (row g, col 2c+1) d2 (row g+8, col 2c)  d3 (row g+8, col 2c+1)
// Rows are windows (g and g+8 of the group), D tile nt holds destination states 8nt + {2c, 2c+1}.
// With the contraction index of k-step ks read as "col c <-> state 8ks + 2c, col c+4 <-> state
// 8ks + 2c + 1", the A fragment of k-step ks is {d0, d2, d1, d3} of D tile nt = ks of the step before.
// ------------------------------------------------------------------------------------------------
template <bool FWD>
__device__ __forceinline__ void b16_chain_run(const B16Args& a, const int grp, const int lane) {
  const int g = lane >> 2, c = lane & 3;
  const int T = a.T, K = a.K, B = a.B;
  const int w0 = grp * 16 + g, w1 = w0 + 8;
  const bool ok0 = w0 < B, ok1 = w1 < B;
  const size_t rb0 = (size_t)(ok0 ? w0 : B - 1) * T, rb1 = (size_t)(ok1 ? w1 : B - 1) * T;
  const float* Pm = FWD ? a.P : a.PT;
  unsigned ph[2][2][2], pl[2][2][2];                      // [k-step][D tile][b0, b1], hi / lo
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int i0 = 8 * ks + 2 * c, i1 = i0 + 1, j = 8 * nt + g;
      const float p0 = (i0 < K && j < K) ? __ldg(Pm + i0 * K + j) : 0.f;
      const float p1 = (i1 < K && j < K) ? __ldg(Pm + i1 * K + j) : 0.f;
      split_tf32(p0, ph[ks][nt][0], pl[ks][nt][0]);
      split_tf32(p1, ph[ks][nt][1], pl[ks][nt][1]);
    }
  const int dt = FWD ? 1 : -1;
  const int tb = FWD ? 0 : T - 1;                         // first row of the recursion
  // this lane's slice of a table row: floats [8nt + 2c, 8nt + 2c + 1] of windows w0 and w1
  const float* bp0 = a.bt + (rb0 + tb) * B16_KS + 2 * c;
  const float* bp1 = a.bt + (rb1 + tb) * B16_KS + 2 * c;
  float* out = FWD ? a.at : a.ct;
  float* op0 = out + (rb0 + tb) * B16_KS + 2 * c;
  float* op1 = out + (rb1 + tb) * B16_KS + 2 * c;
  int* ep0 = a.Et + rb0 + tb; int* ep1 = a.Et + rb1 + tb;
  const int stp = dt * B16_KS;
  // asm volatile: the compiler must leave these loads where they are written (B16_PF steps ahead of
  // their use); as plain __ldg it sank all refills of an unrolled block to the block's end
  auto ldb = [&](const float* p) {
    float2 r;
    asm volatile("ld.global.nc.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
  };
  auto stv = [&](float* p, const float x, const float y, const bool ok) {
    if (ok) *reinterpret_cast<float2*>(p) = make_float2(x, y);
  };
  // ---- first row
  float v[2][4];                                          // recursion vector in D-fragment layout
  {
    const float2 b00 = ldb(bp0), b01 = ldb(bp0 + 8), b10 = ldb(bp1), b11 = ldb(bp1 + 8);
    if (FWD) {
      const int j0 = 2 * c, j1 = 8 + 2 * c;
      const float q00 = j0 < K ? __ldg(a.pi0 + j0) : 0.f, q01 = j0 + 1 < K ? __ldg(a.pi0 + j0 + 1) : 0.f;
      const float q10 = j1 < K ? __ldg(a.pi0 + j1) : 0.f, q11 = j1 + 1 < K ? __ldg(a.pi0 + j1 + 1) : 0.f;
      v[0][0] = q00 * b00.x; v[0][1] = q01 * b00.y; v[0][2] = q00 * b10.x; v[0][3] = q01 * b10.y;
      v[1][0] = q10 * b01.x; v[1][1] = q11 * b01.y; v[1][2] = q10 * b11.x; v[1][3] = q11 * b11.y;
      stv(op0, v[0][0], v[0][1], ok0); stv(op0 + 8, v[1][0], v[1][1], ok0);
      stv(op1, v[0][2], v[0][3], ok1); stv(op1 + 8, v[1][2], v[1][3], ok1);
      if (c == 0) { if (ok0) *ep0 = 0; if (ok1) *ep1 = 0; }
    } else {
      v[0][0] = b00.x; v[0][1] = b00.y; v[0][2] = b10.x; v[0][3] = b10.y;
      v[1][0] = b01.x; v[1][1] = b01.y; v[1][2] = b11.x; v[1][3] = b11.y;
      const int j0 = 2 * c, j1 = 8 + 2 * c;                 // beta[T-1] = 1 on the real states
      const float o00 = j0 < K ? 1.f : 0.f, o01 = j0 + 1 < K ? 1.f : 0.f, o10 = j1 < K ? 1.f : 0.f, o11 = j1 + 1 < K ? 1.f : 0.f;
      stv(op0, o00, o01, ok0); stv(op0 + 8, o10, o11, ok0);
      stv(op1, o00, o01, ok1); stv(op1 + 8, o10, o11, ok1);
    }
  }
  if (T == 1) return;
  // b rows of the next B16_PF steps, kept B16_PF steps ahead of their use
  float2 bq[B16_PF][2][2];
#pragma unroll
  for (int u = 0; u < B16_PF; ++u) {
    const int s = 1 + u;                                   // step index (row tb + s*dt)
    const int so = (s < T ? s : T - 1) * stp;
    bq[u][0][0] = ldb(bp0 + so); bq[u][0][1] = ldb(bp0 + so + 8);
    bq[u][1][0] = ldb(bp1 + so); bq[u][1][1] = ldb(bp1 + so + 8);
  }
  int dprev[2] = {0, 0}, E[2] = {0, 0};
  unsigned xm[2];                                         // row-max bit pattern of the vector two steps back
  xm[0] = xm[1] = (unsigned)B16_TGT << 23;
  auto rowmax = [&](unsigned (&m)[2]) {
    unsigned m0 = max(max(__float_as_uint(v[0][0]), __float_as_uint(v[0][1])), max(__float_as_uint(v[1][0]), __float_as_uint(v[1][1])));
    unsigned m1 = max(max(__float_as_uint(v[0][2]), __float_as_uint(v[0][3])), max(__float_as_uint(v[1][2]), __float_as_uint(v[1][3])));
    m0 = max(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m1 = max(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m0 = max(m0, __shfl_xor_sync(0xffffffffu, m0, 2)); m1 = max(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    m[0] = m0; m[1] = m1;
  };
  // hi/lo split of the carried vector by truncation: hi = top 11 mantissa bits (exact in TF32), lo = v - hi
  // exactly; the tensor core reads the top 11 bits of lo, so v is represented to 2^-22 (the constant
  // matrix P is split with round-to-nearest once, above)
  auto split_trunc = [](const float x, unsigned& hi, unsigned& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
  };
  auto step = [&](const int s, float2 (&bb)[2][2]) {
    // ---- off the dependent chain: exponent shift of this step, scaled b
    int d[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      int dd = (int)(xm[r] >> 23) - B16_TGT - dprev[r];
      if (xm[r] == 0u) dd = 0;                            // all-zero row (cannot happen with pi0, P > 0; padding safety)
      d[r] = max(-60, min(60, dd));
    }
    const float r0 = __uint_as_float(0x3f800000u - ((unsigned)d[0] << 23)), r1 = __uint_as_float(0x3f800000u - ((unsigned)d[1] << 23));
    float br[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      br[nt][0] = bb[0][nt].x * r0; br[nt][1] = bb[0][nt].y * r0;
      br[nt][2] = bb[1][nt].x * r1; br[nt][3] = bb[1][nt].y * r1;
    }
    // ---- on the chain: operand split of v, 12 MMAs (two accumulator chains of 6, small terms first)
    unsigned ah[2][4], al[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      split_trunc(v[ks][0], ah[ks][0], al[ks][0]); split_trunc(v[ks][2], ah[ks][1], al[ks][1]);
      split_trunc(v[ks][1], ah[ks][2], al[ks][2]); split_trunc(v[ks][3], ah[ks][3], al[ks][3]);
    }
    rowmax(xm);                                           // of v[s-1]: consumed by step s+1
    float acc[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      mma_tf32(acc[0], al[ks], ph[ks][0]); mma_tf32(acc[1], al[ks], ph[ks][1]);
      mma_tf32(acc[0], ah[ks], pl[ks][0]); mma_tf32(acc[1], ah[ks], pl[ks][1]);
    }
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) { mma_tf32(acc[0], ah[ks], ph[ks][0]); mma_tf32(acc[1], ah[ks], ph[ks][1]); }
    float* o0 = op0 + s * stp; float* o1 = op1 + s * stp;
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      v[nt][0] = acc[nt][0] * br[nt][0]; v[nt][1] = acc[nt][1] * br[nt][1];
      v[nt][2] = acc[nt][2] * br[nt][2]; v[nt][3] = acc[nt][3] * br[nt][3];
      if (FWD) { stv(o0 + 8 * nt, v[nt][0], v[nt][1], ok0); stv(o1 + 8 * nt, v[nt][2], v[nt][3], ok1); }
      else { stv(o0 + 8 * nt, acc[nt][0] * r0, acc[nt][1] * r0, ok0); stv(o1 + 8 * nt, acc[nt][2] * r1, acc[nt][3] * r1, ok1); }
    }
    E[0] += d[0]; E[1] += d[1];
    dprev[0] = d[0]; dprev[1] = d[1];
    if (FWD && c == 0) { if (ok0) ep0[s] = E[0]; if (ok1) ep1[s] = E[1]; }
    // refill this slot with the row of step s + B16_PF
    const int sn = s + B16_PF;
    const int so = (sn < T ? sn : T - 1) * stp;
    bb[0][0] = ldb(bp0 + so); bb[0][1] = ldb(bp0 + so + 8);
    bb[1][0] = ldb(bp1 + so); bb[1][1] = ldb(bp1 + so + 8);
  };
  int s = 1;
  for (; s + B16_PF <= T; s += B16_PF) {
#pragma unroll
    for (int u = 0; u < B16_PF; ++u) step(s + u, bq[u]);
  }
#pragma unroll
  for (int u = 0; u < B16_PF; ++u) if (s + u < T) step(s + u, bq[u]);
}

__global__ void __launch_bounds__(128) k_b16_chain(const B16Args a, const int ngroups) {
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;        // (group, direction)
  const int grp = wid >> 1;
  if (grp >= ngroups) return;
  if (wid & 1) b16_chain_run<false>(a, grp, lane);
  else b16_chain_run<true>(a, grp, lane);
}

// ------------------------------------------------------------------------------------------------
// phase C: marginals + statistics.  A warp takes units of 8 consecutive rows of one window (+ the row
// that follows them as the "next" side of the last transition pair: the first row of the next unit,
// or row 0 of the window for the reference's wrap-around pair, quirk Q2).
// ------------------------------------------------------------------------------------------------
#define B16_QS 20            // floats per row of the per-warp q tile (16 + 4: conflict-free fragment reads)
#define B16_FLUSH 32         // units between float32 -> float64 flushes of the accumulators

// hi/lo split by truncation (v = hi + lo exactly, the tensor core keeps the top 11 bits of lo): 2
// instructions per value against ~8 for two round-to-nearest conversions; the operands of the
// statistics are sums of thousands of terms, a 2^-22 relative truncation is far below their 1e-5 bound
__device__ __forceinline__ void b16_split(const float x, unsigned& hi, unsigned& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

template <int NTE>
__global__ void __launch_bounds__(256) k_b16_post(const B16Args a) {
  __shared__ __align__(16) float qs_all[8][9][B16_QS];
  __shared__ double S[16 * 16 + 16 * (16 + 16 + 1) + 16 + 4];          // CTA sums: A | sx | sxx | n | q0 | tail
  const int T = a.T, K = a.K, D = a.D, B = a.B;
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  const int g = lane >> 2, c = lane & 3;
  float (*qs)[B16_QS] = qs_all[wp];
  constexpr int ND8 = (NTE - 1) / 2;
  constexpr int NS = 16 * 16 + 16 * 33 + 16 + 4;
  for (int i = tid; i < NS; i += 256) S[i] = 0.0;
  __syncthreads();
  const int ngw = (T + 7) >> 3;                                        // units per window
  const int64_t nunits = (int64_t)B * ngw;
  const int64_t wglob = (int64_t)blockIdx.x * 8 + wp, nwarps = (int64_t)gridDim.x * 8;
  float accT[2][4], accE[NTE][4];
  double dT[2][4], dE[NTE][4];
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) { accT[j][e] = 0.f; dT[j][e] = 0.0; }
#pragma unroll
  for (int j = 0; j < NTE; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) { accE[j][e] = 0.f; dE[j][e] = 0.0; }
  float4 q0a = make_float4(0.f, 0.f, 0.f, 0.f);
  double lzs = 0.0, q4s = 0.0;
  int since = 0;
  const bool vec_out = a.var_x_out && K == 16 && ((((uintptr_t)a.var_x_out) & 15) == 0);
  const bool small = nunits < (int64_t)0x7fffffff;
  const bool f32x = a.dtype == SVIHMM_F32;
  for (int64_t u = wglob; u < nunits; u += nwarps) {
    const int w = small ? (int)((unsigned)u / (unsigned)ngw) : (int)(u / ngw);
    const int t0 = (int)(u - (int64_t)w * ngw) * 8;
    const size_t rbase = (size_t)w * T;
    // ---- marginals of rows t0 .. t0+7 (4 lanes per row, one float4 each) and of the next row
    {
      const int t = t0 + g;
      const bool valid = t < T;
      float4 al = make_float4(0.f, 0.f, 0.f, 0.f), be = al;
      if (valid) {
        al = *reinterpret_cast<const float4*>(a.at + (rbase + t) * B16_KS + 4 * c);
        be = *reinterpret_cast<const float4*>(a.ct + (rbase + t) * B16_KS + 4 * c);
      }
      float4 p = make_float4(al.x * be.x, al.y * be.y, al.z * be.z, al.w * be.w);
      float sa = (al.x + al.y) + (al.z + al.w), sp = (p.x + p.y) + (p.z + p.w);
      sa += __shfl_xor_sync(0xffffffffu, sa, 1); sp += __shfl_xor_sync(0xffffffffu, sp, 1);
      sa += __shfl_xor_sync(0xffffffffu, sa, 2); sp += __shfl_xor_sync(0xffffffffu, sp, 2);
      const float inv = valid ? 1.f / sp : 0.f;
      p.x *= inv; p.y *= inv; p.z *= inv; p.w *= inv;
      *reinterpret_cast<float4*>(&qs[g][4 * c]) = p;
      if (valid && a.var_x_out) {
        float* dst = a.var_x_out + (rbase + t) * K + 4 * c;
        if (vec_out) *reinterpret_cast<float4*>(dst) = p;
        else {
          const float pe[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) if (4 * c + e < K) dst[e] = pe[e];
        }
      }
      if (t == 0) { q0a.x += p.x; q0a.y += p.y; q0a.z += p.z; q0a.w += p.w; }
      // per-row log normalisers; the 8 rows of the unit are summed in float32 (|terms| <~ 1e4: the
      // rounding is ~1e-9 of a window's total), windows and the minibatch in float64
      float lz = 0.f, q4 = 0.f;
      if (valid && c == 0) {
        const float lt = fmaf((float)a.Et[rbase + t], 0.69314718056f, __logf(sa));
        const float mxv = (float)a.mx[rbase + t];
        lz = (t == T - 1 ? lt : 0.f) + mxv;
        q4 = fmaf((float)(T - t), mxv, lt);
      }
#pragma unroll
      for (int o = 4; o < 32; o <<= 1) { lz += __shfl_xor_sync(0xffffffffu, lz, o); q4 += __shfl_xor_sync(0xffffffffu, q4, o); }
      if (lane == 0) {
        atomicAdd(a.seq + 2 * (size_t)w, (double)lz); atomicAdd(a.seq + 2 * (size_t)w + 1, (double)q4);
        lzs += (double)lz; q4s += (double)q4;
      }
      // the row after the unit: t0 + 8, or row 0 for the wrap-around pair of the window's last unit
      int tn = t0 + 8;
      if (tn >= T) tn = (a.wrap && t0 + 8 >= T) ? 0 : -1;
      if (T == 1 && !a.wrap) tn = -1;
      __syncwarp();                                         // the zero rows of a partial unit are written first
      if (lane < 4) {
        float4 pn = make_float4(0.f, 0.f, 0.f, 0.f);
        float spn = 0.f;
        if (tn >= 0) {
          const float4 aln = *reinterpret_cast<const float4*>(a.at + (rbase + tn) * B16_KS + 4 * c);
          const float4 ben = *reinterpret_cast<const float4*>(a.ct + (rbase + tn) * B16_KS + 4 * c);
          pn = make_float4(aln.x * ben.x, aln.y * ben.y, aln.z * ben.z, aln.w * ben.w);
          spn = (pn.x + pn.y) + (pn.z + pn.w);
        }
        spn += __shfl_xor_sync(0x0000000fu, spn, 1);
        spn += __shfl_xor_sync(0x0000000fu, spn, 2);
        const float invn = tn >= 0 ? 1.f / spn : 0.f;
        pn.x *= invn; pn.y *= invn; pn.z *= invn; pn.w *= invn;
        // the pair (last valid row of the unit -> tn) sits at local index (last valid row + 1)
        const int nloc = min(8, T - t0);
        *reinterpret_cast<float4*>(&qs[nloc][4 * c]) = pn;
      }
    }
    __syncwarp();
    // ---- A fragments: Q^T of local rows c and c+4 (states g and g+8), split hi/lo
    unsigned ah[4], al_[4];
    b16_split(qs[c][g], ah[0], al_[0]); b16_split(qs[c][g + 8], ah[1], al_[1]);
    b16_split(qs[c + 4][g], ah[2], al_[2]); b16_split(qs[c + 4][g + 8], ah[3], al_[3]);
    // ---- transition pairs (local row k -> k+1), k = c and c+4; rows past the unit hold zeros
    {
      const int nloc = min(8, T - t0);                     // rows nloc+1 .. 8 of the tile are stale: mask them
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const float x0 = (c + 1 <= nloc) ? qs[c + 1][8 * nt + g] : 0.f;
        const float x1 = (c + 5 <= nloc) ? qs[c + 5][8 * nt + g] : 0.f;
        unsigned bh[2], bl[2];
        b16_split(x0, bh[0], bl[0]); b16_split(x1, bh[1], bl[1]);
        mma_tf32(accT[nt], al_, bh); mma_tf32(accT[nt], ah, bl); mma_tf32(accT[nt], ah, bh);
      }
    }
    // ---- emission features of local rows c and c+4: [x | x^2 | 1], dropped rows (mask / NaN) are zero
    {
      const int ta = t0 + c, tb = t0 + c + 4;
      const bool va = ta < T, vb = tb < T;
      const int64_t ga = a.starts[w] + (va ? ta : 0), gb = a.starts[w] + (vb ? tb : 0);
      float xa[ND8], xb[ND8];
      bool na = false, nb = false;
#pragma unroll
      for (int j = 0; j < ND8; ++j) {
        const int d = 8 * j + g;
        xa[j] = (va && d < D) ? (f32x ? __ldg((const float*)a.obs + ga * D + d) : (float)__ldg((const double*)a.obs + ga * D + d)) : 0.f;
        xb[j] = (vb && d < D) ? (f32x ? __ldg((const float*)a.obs + gb * D + d) : (float)__ldg((const double*)a.obs + gb * D + d)) : 0.f;
        na |= isnan(xa[j]); nb |= isnan(xb[j]);
      }
      const unsigned rowbits = 0x11111111u << c;            // the 8 lanes that hold columns of the same two rows
      // (the ballots are taken by all lanes before any short-circuit: rows past the window end differ by lane)
      const unsigned nana = __ballot_sync(0xffffffffu, na), nanb = __ballot_sync(0xffffffffu, nb);
      const bool dropa = !va || (nana & rowbits) || (a.mask && a.mask[ga]);
      const bool dropb = !vb || (nanb & rowbits) || (a.mask && a.mask[gb]);
#pragma unroll
      for (int j = 0; j < ND8; ++j) {
        const float x0 = dropa ? 0.f : xa[j], x1 = dropb ? 0.f : xb[j];
        unsigned bh[2], bl[2];
        b16_split(x0, bh[0], bl[0]); b16_split(x1, bh[1], bl[1]);
        mma_tf32(accE[j], al_, bh); mma_tf32(accE[j], ah, bl); mma_tf32(accE[j], ah, bh);
        b16_split(x0 * x0, bh[0], bl[0]); b16_split(x1 * x1, bh[1], bl[1]);
        mma_tf32(accE[ND8 + j], al_, bh); mma_tf32(accE[ND8 + j], ah, bl); mma_tf32(accE[ND8 + j], ah, bh);
      }
      unsigned bo[2];
      bo[0] = (!dropa && g == 0) ? 0x3f800000u : 0u;        // the count column: w = 1.0 (exact in TF32)
      bo[1] = (!dropb && g == 0) ? 0x3f800000u : 0u;
      mma_tf32(accE[NTE - 1], al_, bo); mma_tf32(accE[NTE - 1], ah, bo);
    }
    __syncwarp();                                           // the tile is rewritten by the next unit
    if (++since == B16_FLUSH) {
      since = 0;
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) { dT[j][e] += (double)accT[j][e]; accT[j][e] = 0.f; }
#pragma unroll
      for (int j = 0; j < NTE; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) { dE[j][e] += (double)accE[j][e]; accE[j][e] = 0.f; }
    }
  }
  // ---- warp -> CTA (shared-memory float64 atomics) -> global (float64 atomics)
  // fragment element e of tile j: state m = g + 8*(e >> 1), column n = 8j + 2c + (e & 1)
  double* SA = S; double* SX = S + 256; double* SXX = SX + 256; double* SN = SXX + 256; double* SQ0 = SN + 16; double* ST = SQ0 + 16;
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int m = g + 8 * (e >> 1), n = 8 * j + 2 * c + (e & 1);
      atomicAdd(SA + m * 16 + n, dT[j][e] + (double)accT[j][e]);
    }
#pragma unroll
  for (int j = 0; j < NTE; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int m = g + 8 * (e >> 1);
      const double val = dE[j][e] + (double)accE[j][e];
      if (j < ND8) atomicAdd(SX + m * 16 + 8 * j + 2 * c + (e & 1), val);
      else if (j < 2 * ND8) atomicAdd(SXX + m * 16 + 8 * (j - ND8) + 2 * c + (e & 1), val);
      else if (2 * c + (e & 1) == 0) atomicAdd(SN + m, val);
    }
  if (g == 0) {                                             // lanes 0..3 hold the q[0] sums of states 4c .. 4c+3
    atomicAdd(SQ0 + 4 * c, (double)q0a.x); atomicAdd(SQ0 + 4 * c + 1, (double)q0a.y);
    atomicAdd(SQ0 + 4 * c + 2, (double)q0a.z); atomicAdd(SQ0 + 4 * c + 3, (double)q0a.w);
  }
  if (lane == 0) { atomicAdd(ST, lzs); atomicAdd(ST + 1, q4s); }
  __syncthreads();
  const bool first = blockIdx.x == 0;
  for (int e = tid; e < K * K; e += 256) {
    const int i = e / K, j = e - i * K;
    double tot = SA[i * 16 + j];
    if (first && a.add_prior) tot += (double)B * (a.prior_tran[e] - 1.0);
    atomicAdd(a.stats_out + e, tot);
  }
  for (int e = tid; e < K * D; e += 256) {
    const int k = e / D, d = e - k * D;
    atomicAdd(a.stats_out + a.o_sx + e, SX[k * 16 + d]);
    atomicAdd(a.stats_out + a.o_sxx + e, SXX[k * 16 + d]);
  }
  if (tid < K) { atomicAdd(a.stats_out + a.o_n + tid, SN[tid]); atomicAdd(a.stats_out + a.o_q0 + tid, SQ0[tid]); }
  if (tid == 0) {
    atomicAdd(a.stats_out + a.o_tail, ST[0]);
    atomicAdd(a.stats_out + a.o_tail + 1, ST[1]);
    if (first) atomicAdd(a.stats_out + a.o_tail + 2, (double)B);
  }
}
