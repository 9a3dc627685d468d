// Block-parallel scan for LONG chains with K <= 16 (full_local_update / pred_logprob_full /
// hmmbatchcd / hmmbatchsgd at T_full ~ 1e6: hmmsgd_metaobs.py:1147-1205, hmmbase.py:266-320).  The
// sequential kernels advance ONE chain one step at a time (T dependent matvecs on one warp); here
// the T rows are cut into chunks of Lc rows and the recursions become three parallel phases:
//   1  k_scan_ops      per chunk and direction the K x K TRANSFER OPERATOR of the chunk: the product
//                      prod_t (P diag(b_t)) applied to the 16 unit vectors at once = the 16-row
//                      tensor-core recursion of batch16.cuh with rows = start states instead of windows
//                      (mma.m16n8k8, 3xTF32, power-of-two rescaling per row); one warp per chunk
//   2  k_scan_combine  the normalised messages at the chunk boundaries, one small matvec per chunk,
//                      sequentially (one warp per window and direction)
//   3  k_scan_pass     per chunk the ordinary vector recursion from its now known boundary message;
//                      16 chunks per warp as the 16 rows of the same tensor-core step.  Forward: the
//                      normalised alpha-hat and the scale factors c_t; backward: beta-hat on the fly,
//                      q = norm(alpha-hat * beta-hat) (hmmsgd_metaobs.py:516-519) and optionally the
//                      beta-hat / d_t tables
// The outputs are exactly those of k_forward / k_backward (fb.cuh), so everything downstream (log
// normalisers, statistics, svihmm_get_locals) is unchanged.  Work: (K + 1) x the sequential
// recursion, spread over ~T / Lc warps instead of one.
#pragma once
#include "batch16.cuh"

struct ScanArgs {
  int B, T, K, Lc, C;          // C = ceil(T / Lc) chunks per window
  const float *P, *PT, *pi0;
  const float* b;              // (B*T, K) scaled likelihoods
  float* alpha; float* cs;     // (B*T, K), (B*T)
  float* q; float* beta; float* sb;
  float* ops;                  // [B*C][2][272]: 16 x 16 operator (row-major, scaled) + 16 row exponents
  float* bound;                // [B*C][2][16]: forward message at the END of the chunk / backward message at its LAST row
};
#define SC_OP 272

struct Sc16P { unsigned ph[2][2][2], pl[2][2][2]; };

__device__ __forceinline__ void sc16_load_P(Sc16P& f, const float* __restrict__ Pm, const int K, const int g, const int c) {
#pragma unroll
  for (int ks = 0; ks < 2; ++ks)
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int i0 = 8 * ks + 2 * c, i1 = i0 + 1, j = 8 * nt + g;
      const float p0 = (i0 < K && j < K) ? __ldg(Pm + i0 * K + j) : 0.f;
      const float p1 = (i1 < K && j < K) ? __ldg(Pm + i1 * K + j) : 0.f;
      split_tf32(p0, f.ph[ks][nt][0], f.pl[ks][nt][0]);
      split_tf32(p1, f.ph[ks][nt][1], f.pl[ks][nt][1]);
    }
}
// acc = v . P  (v in D-fragment layout of the previous step, see batch16.cuh)
__device__ __forceinline__ void sc16_mma(const float (&v)[2][4], const Sc16P& f, float (&acc)[2][4]) {
  unsigned ah[2][4], al[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    b16_split(v[ks][0], ah[ks][0], al[ks][0]); b16_split(v[ks][2], ah[ks][1], al[ks][1]);
    b16_split(v[ks][1], ah[ks][2], al[ks][2]); b16_split(v[ks][3], ah[ks][3], al[ks][3]);
  }
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    mma_tf32(acc[0], al[ks], f.ph[ks][0]); mma_tf32(acc[1], al[ks], f.ph[ks][1]);
    mma_tf32(acc[0], ah[ks], f.pl[ks][0]); mma_tf32(acc[1], ah[ks], f.pl[ks][1]);
  }
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) { mma_tf32(acc[0], ah[ks], f.ph[ks][0]); mma_tf32(acc[1], ah[ks], f.ph[ks][1]); }
}
// sums over the 16 columns of rows g (x[.][0..1]) and g+8 (x[.][2..3]): quad reduction
__device__ __forceinline__ void sc16_rowsum(const float (&x)[2][4], float& s0, float& s1) {
  s0 = (x[0][0] + x[0][1]) + (x[1][0] + x[1][1]);
  s1 = (x[0][2] + x[0][3]) + (x[1][2] + x[1][3]);
  s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
  s0 += __shfl_xor_sync(0xffffffffu, s0, 2); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
}
// exponent shift that brings a row whose sum is `s` back to 2^40 (0 for an all-zero padding row)
__device__ __forceinline__ int sc16_shift(const float s) {
  const unsigned u = __float_as_uint(s);
  if (u == 0u) return 0;
  return max(-100, min(100, (int)(u >> 23) - B16_TGT));
}
__device__ __forceinline__ float sc16_pow2(const int d) { return __uint_as_float((unsigned)(127 - d) << 23); }

// ---- phase 1: transfer operators.  One warp per (window, chunk, direction); lane (g, c) holds rows g
// and g+8 of the 16 x 16 operator in D-fragment layout.  The rescaling uses the row sum of the step
// before (one step of lag is harmless here: no store depends on it).
__global__ void __launch_bounds__(128) k_scan_ops(const ScanArgs a) {
  const int lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nch = a.B * a.C;
  if (wid >= 2 * nch) return;
  const bool fwd = !(wid & 1);
  const int gc = wid >> 1, w = gc / a.C, ch = gc - w * a.C;
  const int K = a.K, T = a.T;
  const int t0 = ch * a.Lc, t1 = min(T, t0 + a.Lc);
  Sc16P f;
  sc16_load_P(f, fwd ? a.P : a.PT, K, g, c);
  const float* bw = a.b + (size_t)w * T * K;
  auto ldrow = [&](const int t, float (&bj)[2][2]) {      // b[t][8nt + 2c], b[t][8nt + 2c + 1]
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int j = 8 * nt + 2 * c;
      bj[nt][0] = j < K ? __ldg(bw + (size_t)t * K + j) : 0.f;
      bj[nt][1] = j + 1 < K ? __ldg(bw + (size_t)t * K + j + 1) : 0.f;
    }
  };
  float v[2][4];
  int E0 = 0, E1 = 0;
  // identity rows (only real states), or the identity times the first row's b
  const bool withb = !fwd || ch == 0;
  {
    float bj[2][2];
    ldrow(fwd ? t0 : t1 - 1, bj);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int j0 = 8 * nt + 2 * c, j1 = j0 + 1;
      v[nt][0] = (g == j0 && j0 < K) ? (withb ? bj[nt][0] : 1.f) : 0.f;
      v[nt][1] = (g == j1 && j1 < K) ? (withb ? bj[nt][1] : 1.f) : 0.f;
      v[nt][2] = (g + 8 == j0 && j0 < K) ? (withb ? bj[nt][0] : 1.f) : 0.f;
      v[nt][3] = (g + 8 == j1 && j1 < K) ? (withb ? bj[nt][1] : 1.f) : 0.f;
    }
  }
  // rows still to process: forward t = first..t1-1 ascending; backward t = t1-2..t0 descending
  const int first = fwd ? (ch == 0 ? t0 + 1 : t0) : t1 - 2;
  const int nstep = fwd ? t1 - first : first - t0 + 1;
  const int dt = fwd ? 1 : -1;
  float bq[4][2][2];
#pragma unroll
  for (int u = 0; u < 4; ++u) ldrow(min(max(first + u * dt, 0), T - 1), bq[u]);
  float s0 = 1.f, s1 = 1.f;
  for (int s = 0; s < nstep; s += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (s + u < nstep) {
        const int d0 = sc16_shift(s0), d1 = sc16_shift(s1);
        const float r0 = sc16_pow2(d0), r1 = sc16_pow2(d1);
        float acc[2][4];
        sc16_mma(v, f, acc);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          v[nt][0] = acc[nt][0] * (bq[u][nt][0] * r0); v[nt][1] = acc[nt][1] * (bq[u][nt][1] * r0);
          v[nt][2] = acc[nt][2] * (bq[u][nt][0] * r1); v[nt][3] = acc[nt][3] * (bq[u][nt][1] * r1);
        }
        E0 += d0; E1 += d1;
        sc16_rowsum(v, s0, s1);
        ldrow(min(max(first + (s + u + 4) * dt, 0), T - 1), bq[u]);
      }
    }
  }
  float* op = a.ops + ((size_t)gc * 2 + (fwd ? 0 : 1)) * SC_OP;
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    const int j = 8 * nt + 2 * c;
    op[g * 16 + j] = v[nt][0]; op[g * 16 + j + 1] = v[nt][1];
    op[(g + 8) * 16 + j] = v[nt][2]; op[(g + 8) * 16 + j + 1] = v[nt][3];
  }
  if (c == 0) { op[256 + g] = __int_as_float(E0); op[256 + g + 8] = __int_as_float(E1); }
}

// ---- phase 2: boundary messages.  One warp per (window, direction); lane j < 16 holds component j.
//   forward   a_c = norm( sum_i a_{c-1}[i] 2^E_i Phi_c[i][:] ),  a_{-1} = pi0 (chunk 0's operator starts
//             with diag(b_0) instead of a transition)
//   backward  h_{C-1} = 1;  h_{c-1} = norm( P u ),  u = sum_j h_c[j] 2^E_j V_c[j][:]
__global__ void __launch_bounds__(64) k_scan_combine(const ScanArgs a) {
  const int lane = threadIdx.x & 31, dir = threadIdx.x >> 5, w = blockIdx.x;
  const int K = a.K, C = a.C;
  const int j = lane & 15;
  const bool act = lane < 16 && j < K;
  float Prow[16];                                          // backward: row j of P
#pragma unroll
  for (int i = 0; i < 16; ++i) Prow[i] = (dir == 1 && act && i < K) ? __ldg(a.P + j * K + i) : 0.f;
  float x = dir == 0 ? (act ? __ldg(a.pi0 + j) : 0.f) : (act ? 1.f : 0.f);
  if (dir == 1 && lane < 16) a.bound[(((size_t)w * C + (C - 1)) * 2 + 1) * 16 + j] = x;
  const int cbeg = dir == 0 ? 0 : C - 1, cend = dir == 0 ? C : 0, dc = dir == 0 ? 1 : -1;
  // the operator of the NEXT chunk is loaded while the current one is applied (the chain of matvecs
  // would otherwise wait for an L2 round trip per chunk)
  float opn[16]; int En;
  auto load_op = [&](const int ch) {
    const float* op = a.ops + (((size_t)w * C + ch) * 2 + dir) * SC_OP;
#pragma unroll
    for (int i = 0; i < 16; ++i) opn[i] = lane < 16 ? __ldg(op + i * 16 + j) : 0.f;
    En = act ? __float_as_int(__ldg(op + 256 + j)) : -100000;     // padding rows (i >= K) carry no scale
  };
  load_op(cbeg);
  for (int ch = cbeg; ch != cend; ch += dc) {
    float opc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) opc[i] = opn[i];
    const int Ei = En;
    if (ch + dc != cend) load_op(ch + dc);
    int Em = Ei;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) Em = max(Em, __shfl_xor_sync(0xffffffffu, Em, o));
    Em = __shfl_sync(0xffffffffu, Em, 0);
    // weight of operator row i (held by lane i): x[i] * 2^(E_i - Emax); rows far below the largest vanish
    const int de = max(Ei - Em, -120);
    const float wi = lane < 16 ? x * __uint_as_float((unsigned)(127 + de) << 23) : 0.f;
    float y = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) y = fmaf(__shfl_sync(0xffffffffu, wi, i), opc[i], y);
    if (dir == 1) {                                        // h = P u
      float z = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) z = fmaf(Prow[i], __shfl_sync(0xffffffffu, y, i), z);
      y = z;
    }
    float sum = lane < 16 ? y : 0.f;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    sum = __shfl_sync(0xffffffffu, sum, 0);
    x = y * (1.f / sum);
    if (dir == 0) { if (lane < 16) a.bound[(((size_t)w * C + ch) * 2) * 16 + j] = x; }
    else if (ch > 0 && lane < 16) a.bound[(((size_t)w * C + ch - 1) * 2 + 1) * 16 + j] = x;
  }
}

// ---- phase 3: vector passes, 16 chunks per warp.  Row r of the warp = global chunk grp*16 + r.
template <bool FWD>
__device__ __forceinline__ void sc16_pass(const ScanArgs& a, const int grp, const int lane) {
  const int g = lane >> 2, c = lane & 3;
  const int K = a.K, T = a.T, C = a.C, Lc = a.Lc;
  const int nch = a.B * C;
  int gcr[2], t0r[2], t1r[2]; bool okr[2]; size_t wb[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int gc = grp * 16 + g + 8 * r;
    okr[r] = gc < nch;
    gcr[r] = okr[r] ? gc : nch - 1;
    const int w = gcr[r] / C, ch = gcr[r] - w * C;
    t0r[r] = ch * Lc; t1r[r] = min(T, t0r[r] + Lc);
    wb[r] = (size_t)w * T;
  }
  Sc16P f;
  sc16_load_P(f, FWD ? a.P : a.PT, K, g, c);
  const int j00 = 2 * c, j10 = 8 + 2 * c;                  // this lane's columns: j00, j00+1, j10, j10+1
  auto ld4 = [&](const float* tab, const size_t row, float (&o)[4]) {
    const float* p = tab + row * K;
    o[0] = j00 < K ? __ldg(p + j00) : 0.f; o[1] = j00 + 1 < K ? __ldg(p + j00 + 1) : 0.f;
    o[2] = j10 < K ? __ldg(p + j10) : 0.f; o[3] = j10 + 1 < K ? __ldg(p + j10 + 1) : 0.f;
  };
  auto st4 = [&](float* tab, const size_t row, const float x0, const float x1, const float x2, const float x3) {
    float* p = tab + row * K;
    if (j00 < K) p[j00] = x0; if (j00 + 1 < K) p[j00 + 1] = x1;
    if (j10 < K) p[j10] = x2; if (j10 + 1 < K) p[j10 + 1] = x3;
  };
  // boundary message each row starts from: forward a_{ch-1} (pi0 for chunk 0), backward h_ch
  float v[2][4];
  {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int ch = gcr[r] % C;
      float bd[4];
      if (FWD) {
        if (ch == 0) { bd[0] = j00 < K ? __ldg(a.pi0 + j00) : 0.f; bd[1] = j00 + 1 < K ? __ldg(a.pi0 + j00 + 1) : 0.f;
                       bd[2] = j10 < K ? __ldg(a.pi0 + j10) : 0.f; bd[3] = j10 + 1 < K ? __ldg(a.pi0 + j10 + 1) : 0.f; }
        else { const float* p = a.bound + ((size_t)(gcr[r] - 1) * 2) * 16; bd[0] = p[j00]; bd[1] = p[j00 + 1]; bd[2] = p[j10]; bd[3] = p[j10 + 1]; }
      } else {
        const float* p = a.bound + ((size_t)gcr[r] * 2 + 1) * 16; bd[0] = p[j00]; bd[1] = p[j00 + 1]; bd[2] = p[j10]; bd[3] = p[j10 + 1];
      }
      v[0][2 * r] = bd[0]; v[0][2 * r + 1] = bd[1]; v[1][2 * r] = bd[2]; v[1][2 * r + 1] = bd[3];
    }
  }
  // Step s handles row t = t0 + s (forward) / t1 - 1 - s (backward) of every chunk.  m = "message before
  // the likelihood of row t": the boundary message itself at s = 0 of the first chunk (forward: pi0) or
  // of every chunk (backward: h_ch), else v . P.  Then v = m * b[t] * 2^-d.
  float sprev[2] = {1.f, 1.f};                             // forward: sum of v after the previous step; backward: sum of m
  float rprev[2] = {1.f, 1.f};
  for (int s = 0; s <= Lc; ++s) {
    float acc[2][4];
    sc16_mma(v, f, acc);
    int tr[2]; bool live[2], bndry[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      tr[r] = FWD ? t0r[r] + s : t1r[r] - 1 - s;
      live[r] = okr[r] && (FWD ? tr[r] < t1r[r] : tr[r] >= t0r[r]);
      bndry[r] = s == 0 && (FWD ? t0r[r] == 0 : true);    // m is the boundary message, not a product
    }
    // backward only: one more product past the chunk's first row gives the scale factor of the row
    // before it (the last row of the previous chunk): d = sum(P (beta-hat b)) with beta-hat normalised
    float m[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        m[nt][2 * r] = bndry[r] ? v[nt][2 * r] : acc[nt][2 * r];
        m[nt][2 * r + 1] = bndry[r] ? v[nt][2 * r + 1] : acc[nt][2 * r + 1];
      }
    float ms0, ms1;
    sc16_rowsum(m, ms0, ms1);
    const float msum[2] = {ms0, ms1};
    if (!FWD) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        // row tr: beta-hat = m / sum(m); its scale factor d_t = sum(m) / (rprev * sprev) (fb.cuh: S1)
        const bool past = okr[r] && tr[r] == t0r[r] - 1 && tr[r] >= 0;       // one row past the chunk: only d_t
        const bool lastrow = bndry[r] && (gcr[r] % C) == C - 1;              // row T-1: d = 1 by definition
        // (the d_t of a chunk's last row is written by the NEXT chunk's "past" step, not here)
        if (a.sb && c == 0 && okr[r] && ((live[r] && !bndry[r]) || past || lastrow))
          a.sb[wb[r] + tr[r]] = lastrow ? 1.f : msum[r] / (rprev[r] * sprev[r]);
      }
    }
    if (s == Lc) break;
    float bb[2][4], aa[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const size_t row = wb[r] + (size_t)min(max(tr[r], 0), T - 1);
      ld4(a.b, row, bb[r]);
      if (!FWD) ld4(a.alpha, row, aa[r]);
    }
    int d[2]; float rr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) { d[r] = sc16_shift(msum[r]); rr[r] = sc16_pow2(d[r]); }
    if (!FWD) {
      // marginals of row tr: q = alpha-hat * m / sum
      float p[2][4];
#pragma unroll
      for (int r = 0; r < 2; ++r) { p[0][2 * r] = aa[r][0] * m[0][2 * r]; p[0][2 * r + 1] = aa[r][1] * m[0][2 * r + 1];
                                    p[1][2 * r] = aa[r][2] * m[1][2 * r]; p[1][2 * r + 1] = aa[r][3] * m[1][2 * r + 1]; }
      float ps0, ps1;
      sc16_rowsum(p, ps0, ps1);
      // beta-hat[T-1] = 1 on every state, NOT normalised (lbeta[T-1] = 0, hmmsgd_metaobs.py:850; fb.cuh)
      const bool last0 = bndry[0] && (gcr[0] % C) == C - 1, last1 = bndry[1] && (gcr[1] % C) == C - 1;
      const float pinv[2] = {1.f / ps0, 1.f / ps1}, minv[2] = {last0 ? 1.f : 1.f / ms0, last1 ? 1.f : 1.f / ms1};
#pragma unroll
      for (int r = 0; r < 2; ++r)
        if (live[r]) {
          const size_t row = wb[r] + tr[r];
          st4(a.q, row, p[0][2 * r] * pinv[r], p[0][2 * r + 1] * pinv[r], p[1][2 * r] * pinv[r], p[1][2 * r + 1] * pinv[r]);
          if (a.beta) st4(a.beta, row, m[0][2 * r] * minv[r], m[0][2 * r + 1] * minv[r], m[1][2 * r] * minv[r], m[1][2 * r + 1] * minv[r]);
        }
    }
    // v = m * b * 2^-d
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      v[0][2 * r] = m[0][2 * r] * (bb[r][0] * rr[r]); v[0][2 * r + 1] = m[0][2 * r + 1] * (bb[r][1] * rr[r]);
      v[1][2 * r] = m[1][2 * r] * (bb[r][2] * rr[r]); v[1][2 * r + 1] = m[1][2 * r + 1] * (bb[r][3] * rr[r]);
    }
    if (FWD) {
      float vs0, vs1;
      sc16_rowsum(v, vs0, vs1);
      const float vsum[2] = {vs0, vs1};
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (live[r]) {
          const size_t row = wb[r] + tr[r];
          const float inv = 1.f / vsum[r];
          st4(a.alpha, row, v[0][2 * r] * inv, v[0][2 * r + 1] * inv, v[1][2 * r] * inv, v[1][2 * r + 1] * inv);
          // c_t = sum_j (alpha-hat_{t-1} P)_j b_t(j): the sum now over the sum before, undoing this step's 2^-d
          if (c == 0) a.cs[row] = vsum[r] / (rr[r] * sprev[r]);        // sprev = 1 at a chunk's first row (normalised boundary)
        }
        sprev[r] = vsum[r];
      }
    } else {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const bool lastr = bndry[r] && (gcr[r] % C) == C - 1;      // row T-1: the all-ones vector counts as "normalised"
        sprev[r] = lastr ? 1.f : msum[r]; rprev[r] = rr[r];
      }
    }
  }
}

__global__ void __launch_bounds__(128) k_scan_pass(const ScanArgs a, const int ngroups, const int fwd) {
  const int lane = threadIdx.x & 31;
  const int grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (grp >= ngroups) return;
  if (fwd) sc16_pass<true>(a, grp, lane);
  else sc16_pass<false>(a, grp, lane);
}
