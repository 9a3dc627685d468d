// Fused E-step for K <= 32: ONE kernel launch per minibatch, one CTA per window, everything
// between the coalesced read of the window and the write of the posterior marginals stays in
// shared memory.  Replaces, per window, the whole body of the reference's per-meta-observation
// loop hmmsgd_metaobs.py:405-436:
//   phase A  expected log-likelihoods (float64; pybasicbayes/distributions.py:351-366 via the
//            constants of prep.cuh) + np.nan_to_num semantics (hmmsgd_metaobs.py:508-509),
//            b[t][k] = exp(ll - max_k ll) (float32), row maxima (float64); one thread per row,
//            so the row maximum needs no cross-lane traffic
//   phase B  forward (hmmsgd_metaobs.py:775-803) and backward (:828-855) recursions, run
//            CONCURRENTLY by the two lane groups of one warp (the two chains are independent).
//            Only the matvec is on the dependent chain: the K-vector is exchanged through a
//            double-buffered shared-memory slot (STS + 128-bit broadcast loads: ~30 cycles measured,
//            against ~100 for 16 shuffles), the dot product runs on packed fma.rn.f32x2, and
//            instead of a per-step normaliser the messages are rescaled by exact powers of two
//            chosen from the group-max exponent seen two steps earlier (deadbeat controller), so
//            no reduction sits on the critical path and the rescaling adds no rounding error
//   phase C  marginals q[t] = norm(alpha[t]*beta[t]) (:516-519), log normalisers (:257-271),
//            transition statistic sum_t outer(q[t-1],q[t]) with the reference's wrap-around
//            (:876-878) and the weighted NIW statistics (util.py:73-83) as register-blocked 4x4
//            tiles split over row chunks, accumulated over the minibatch (:430-433) with float64
//            atomics into the packed statistics buffer.
//
// Shared memory: [ b | beta | alpha->q ] each T*KS floats (KS = K rounded up to 4) + per-row
// scalars.  The window's observations are staged (float64 for phase A, float32 for phase C) into
// regions that are dead at that point, so HBM traffic is the algorithmic minimum: T*D reads and
// T*K writes per window.
#pragma once
#include "common.cuh"

#define FUSED_NT 256
#define FUSED_XTB 157          // target biased exponent of the group max: 2^30
#define FUSED_RED_BYTES (FUSED_NT * 32 * 4)

struct FusedArgs {
  int B, T, K, D, DD, diag, wrap, add_prior, mask_ll;
  int tri;                   // D(D+1)/2 (full) or D (diag): emission params per state besides gk
  const void* obs; int dtype; const uint8_t* mask; const int64_t* starts;
  const float* Pt; const float* pi0;
  const double* Rs; const double* gk; const double* ck; const double* prior_tran;
  float* var_x_out; double* stats_out; double* seq;
  size_t o_n, o_sx, o_sxx, o_q0, o_tail;
  long long* dbg;            // optional [B][8] clock64 stamps at phase boundaries (debug)
};

struct FusedSmem {           // byte offsets into the dynamic shared memory
  int KS, DS, XP;
  size_t b, c, a, xs, xf, red, mx, late, flags, bc, total;
};

__host__ __device__ inline size_t fused_al16(size_t v) { return (v + 15) & ~(size_t)15; }

__host__ __device__ inline FusedSmem fused_smem_layout(int T, int K, int D, int tri, int diag) {
  FusedSmem s;
  s.KS = (K + 3) & ~3;
  s.DS = (D + 1 + 3) & ~3;                          // float columns of phase C: x_0..x_{D-1}, w, pad
  s.XP = D | 1;                                     // double columns of phase A (odd: conflict-free)
  const size_t R = (size_t)T * s.KS * sizeof(float);
  const size_t xf_bytes = fused_al16((size_t)T * s.DS * sizeof(float));
  const size_t scratchC = xf_bytes + FUSED_RED_BYTES;
  s.b = 0; s.c = R;
  s.a = 2 * R > scratchC ? 2 * R : scratchC;
  s.xs = s.c;                                       // phase A staging over [beta | alpha]
  s.xf = 0; s.red = xf_bytes;                       // phase C scratch over [b | beta]
  size_t end = s.a + R;
  const size_t endA = s.xs + (size_t)T * s.XP * 8;
  if (endA > end) end = endA;
  s.mx = fused_al16(end);                           // T doubles: row maxima of ll
  s.late = fused_al16(s.mx + (size_t)T * 8);        // phase A: emission constants; B/C: lt (T doubles) + E (T ints)
  const size_t params = diag ? ((size_t)s.KS * D * 16 + (size_t)s.KS * 8) : ((size_t)K * (tri + D + 1) * 8);
  const size_t late = (size_t)T * 12;
  s.flags = s.late + fused_al16(params > late ? params : late);
  s.bc = fused_al16(s.flags + (size_t)T);           // 2 parities x 2 warps x 32 floats
  s.total = s.bc + 2 * 2 * 32 * sizeof(float);
  return s;
}

__device__ __forceinline__ void ffma2(unsigned long long& acc, const unsigned long long a, const unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ unsigned long long pack2(const float x, const float y) {
  return (unsigned long long)__float_as_uint(x) | ((unsigned long long)__float_as_uint(y) << 32);
}
__device__ __forceinline__ float lo32(const unsigned long long v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi32(const unsigned long long v) { return __uint_as_float((unsigned)(v >> 32)); }

// One step of either chain (see the header comment).  v: this lane's component of the recursion
// vector (alpha~ for the forward group, b*beta~ for the backward group).  With x_s the biased
// exponent of max_j v_s[j] and d_s the exponent shift applied at step s,
//     d_s = (x_{s-2} - XTB) - d_{s-1}      =>      x_s = XTB + g_{s-1} + g_s
// (g = log2 of the per-step growth, bounded): the shift already in flight is subtracted, so the
// two-step-old measurement keeps the exponent bounded while the max is computed off the chain.
template <int KP>
__device__ __forceinline__ void chain_step(float& v, const unsigned long long (&col2)[KP / 2], const float bt,
                                           const bool fwd, const bool st, const float* bcr, float* bcw_next,
                                           float* op_prev, int* ep_prev, const bool lead, float& pend, int& xa,
                                           int& da, int& E) {
  constexpr int NV = KP / 4;
  __syncwarp();                                        // v of the previous step is in the slot
  float4 x[NV];
#pragma unroll
  for (int q = 0; q < NV; ++q) x[q] = reinterpret_cast<const float4*>(bcr)[q];
  // side traffic of the PREVIOUS step, queued behind the loads the chain is waiting for (measured:
  // 114 -> 101 cycles per step against issuing it right after the broadcast store)
  if (st) *op_prev = pend;
  if (lead) *ep_prev = E;
  int d = xa - FUSED_XTB - da;
  d = max(-60, min(60, d));
  const float r = __uint_as_float((unsigned)(127 - d) << 23);
  const float br = bt * r;
  unsigned long long acc0 = 0ull, acc1 = 0ull;
  unsigned mx = 0u;
#pragma unroll
  for (int q = 0; q < NV; ++q) {
    ffma2(acc0, pack2(x[q].x, x[q].y), col2[2 * q]);
    ffma2(acc1, pack2(x[q].z, x[q].w), col2[2 * q + 1]);
    mx = max(mx, __vimax3_u32(__float_as_uint(x[q].x), __float_as_uint(x[q].y), __float_as_uint(x[q].z)));
    mx = max(mx, __float_as_uint(x[q].w));
  }
  const float m = (lo32(acc0) + hi32(acc0)) + (lo32(acc1) + hi32(acc1));
  v = m * br;
  *bcw_next = v;                                       // the only store on the dependent chain
  E += d;
  pend = fwd ? v : m * r;
  xa = (int)(mx >> 23); da = d;
}

template <int KP>
__global__ void __launch_bounds__(FUSED_NT, 2) k_estep_fused(const FusedArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int T = a.T, K = a.K, D = a.D;
  const FusedSmem L = fused_smem_layout(T, K, D, a.tri, a.diag);
  const int KS = L.KS, DS = L.DS, XP = L.XP;
  float* bS = reinterpret_cast<float*>(smem + L.b);
  float* cS = reinterpret_cast<float*>(smem + L.c);
  float* aS = reinterpret_cast<float*>(smem + L.a);
  double* xs = reinterpret_cast<double*>(smem + L.xs);
  float* xf = reinterpret_cast<float*>(smem + L.xf);
  float* red = reinterpret_cast<float*>(smem + L.red);
  double* mxS = reinterpret_cast<double*>(smem + L.mx);
  double* parS = reinterpret_cast<double*>(smem + L.late);           // phase A only
  double* ltS = reinterpret_cast<double*>(smem + L.late);            // phase B/C
  int* ES = reinterpret_cast<int*>(smem + L.late + (size_t)T * 8);
  unsigned char* fl = smem + L.flags;
  float* bcS = reinterpret_cast<float*>(smem + L.bc);
  const int tid = threadIdx.x, w = blockIdx.x;
  const int64_t s0 = a.starts[w];

#define FUSED_STAMP(i) do { if (a.dbg && tid == 0) a.dbg[(size_t)w * 8 + (i)] = clock64(); } while (0)
  FUSED_STAMP(0);
  // ---------------------------------------------------------------- phase A: emissions
  {
    // stage the window as float64 rows of XP (odd) doubles: coalesced 128-bit global loads
    const int64_t e0 = s0 * D;
    const int n = T * D;
    if (a.dtype == SVIHMM_F32) {
      const float* src = (const float*)a.obs + e0;
      if ((((uintptr_t)src) & 15) == 0 && (D & 3) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        for (int i = tid; i < n / 4; i += FUSED_NT) {
          const float4 q = __ldg(s4 + i);
          const int r = (4 * i) / D, d = 4 * i - r * D;
          double* dst = xs + (size_t)r * XP + d;
          dst[0] = q.x; dst[1] = q.y; dst[2] = q.z; dst[3] = q.w;
        }
      } else {
        for (int i = tid; i < n; i += FUSED_NT) { const int r = i / D; xs[(size_t)r * XP + (i - r * D)] = __ldg(src + i); }
      }
    } else {
      const double* src = (const double*)a.obs + e0;
      if ((((uintptr_t)src) & 15) == 0 && (D & 1) == 0) {
        const double2* s2 = reinterpret_cast<const double2*>(src);
        for (int i = tid; i < n / 2; i += FUSED_NT) {
          const double2 q = __ldg(s2 + i);
          const int r = (2 * i) / D, d = 2 * i - r * D;
          double* dst = xs + (size_t)r * XP + d;
          dst[0] = q.x; dst[1] = q.y;
        }
      } else {
        for (int i = tid; i < n; i += FUSED_NT) { const int r = i / D; xs[(size_t)r * XP + (i - r * D)] = __ldg(src + i); }
      }
    }
    // emission constants.  diagonal (prepared by the global-step kernel): (c2, c1) = (-Rs, 2 Rs mu) per
    // [d][k] and ck' = ck - sum Rs mu^2, so ll = ck' + sum_d (c2 x^2 + c1 x) (float64: the expansion
    // costs ~1e-12 absolute, two DFMA per term)
    if (a.diag) {
      for (int i = tid; i < KS * D; i += FUSED_NT) {      // [d][KS], zero-padded columns
        const int d = i / KS, kk = i - d * KS;
        const bool in = kk < K;
        parS[2 * i] = in ? a.Rs[2 * (d * K + kk)] : 0.0;
        parS[2 * i + 1] = in ? a.Rs[2 * (d * K + kk) + 1] : 0.0;
      }
      for (int kk = tid; kk < KS; kk += FUSED_NT) parS[2 * KS * D + kk] = kk < K ? a.ck[kk] : 0.0;
    } else {
      const int np = a.tri + D + 1;
      for (int i = tid; i < K * np; i += FUSED_NT) {
        const int kk = i / np, p = i - kk * np;
        parS[i] = p < a.tri ? a.Rs[(size_t)kk * a.tri + p] : (p < a.tri + D ? a.gk[(size_t)kk * D + (p - a.tri)] : a.ck[kk]);
      }
    }
    __syncthreads();
    for (int row = tid; row < T; row += FUSED_NT) {
      const double* x = xs + (size_t)row * XP;
      bool bad = false;
      for (int d = 0; d < D; ++d) bad |= isnan(x[d]);
      const bool mk = a.mask && a.mask[s0 + row];
      const bool noev = bad || (a.mask_ll && mk);
      fl[row] = (unsigned char)(((bad || mk) ? 1 : 0) | (noev ? 2 : 0));
      double ll[KP];
      if (a.diag) {
#pragma unroll
        for (int k = 0; k < KP; ++k) ll[k] = k < KS ? parS[2 * KS * D + k] : 0.0;
        for (int d = 0; d < D; ++d) {                   // independent accumulation chains, 4 loads in flight
          const double xd = x[d], xx = xd * xd;
          const double2* pp = reinterpret_cast<const double2*>(parS) + (size_t)d * KS;
#pragma unroll
          for (int k0 = 0; k0 < KP; k0 += 4) {
            if (k0 < KS) {
              const double2 c0 = pp[k0], c1 = pp[k0 + 1], c2 = pp[k0 + 2], c3 = pp[k0 + 3];
              ll[k0] = fma(c0.x, xx, fma(c0.y, xd, ll[k0]));
              ll[k0 + 1] = fma(c1.x, xx, fma(c1.y, xd, ll[k0 + 1]));
              ll[k0 + 2] = fma(c2.x, xx, fma(c2.y, xd, ll[k0 + 2]));
              ll[k0 + 3] = fma(c3.x, xx, fma(c3.y, xd, ll[k0 + 3]));
            }
          }
        }
      } else {
        const int np = a.tri + D + 1;
#pragma unroll
        for (int k = 0; k < KP; ++k) {
          ll[k] = -INFINITY;
          if (k < K) {
            const double* pr = parS + (size_t)k * np;
            double acc = 0.0;
            int o = 0;
            for (int i = 0; i < D; ++i) {
              double s = -pr[a.tri + i];
              for (int j = 0; j <= i; ++j) s = fma(pr[o + j], x[j], s);
              o += i + 1;
              acc = fma(s, s, acc);
            }
            ll[k] = pr[a.tri + D] - acc;
          }
        }
      }
      double m = -INFINITY;
#pragma unroll
      for (int k = 0; k < KP; ++k) { if (k < K) { if (noev) ll[k] = 0.0; m = fmax(m, ll[k]); } }
      mxS[row] = m;
      float* bp = bS + (size_t)row * KS;
#pragma unroll
      for (int k4 = 0; k4 < KP; k4 += 4) {
        if (k4 < KS) {
          float4 o;
          o.x = k4 < K ? __expf((float)(ll[k4] - m)) : 0.f;
          o.y = k4 + 1 < K ? __expf((float)(ll[k4 + 1] - m)) : 0.f;
          o.z = k4 + 2 < K ? __expf((float)(ll[k4 + 2] - m)) : 0.f;
          o.w = k4 + 3 < K ? __expf((float)(ll[k4 + 3] - m)) : 0.f;
          *reinterpret_cast<float4*>(bp + k4) = o;
        }
      }
    }
  }
  __syncthreads();
  FUSED_STAMP(1);

  // ---------------------------------------------------------------- phase B: the two chains
  {
    constexpr int GPW = 32 / KP;                        // lane groups per warp
    const int lane = tid & 31, wp = tid >> 5;
    const int j = lane % KP, grp = wp * GPW + lane / KP;
    if (wp < (KP == 32 ? 2 : 1)) {                      // whole warps only (__syncwarp inside)
      const bool live = grp < 2;                        // lane groups >= 2 of warp 0 (KP < 16) idle along
      const bool fwd = grp == 0;
      const bool act = live && j < K;
      const bool st = live && j < KS;                   // columns K..KS-1 of the tables are kept at zero
      const bool lead = fwd && j == 0;
      unsigned long long col2[KP / 2];
#pragma unroll
      for (int i = 0; i < KP; i += 2) {
        const float p0 = (act && i < K) ? (fwd ? __ldg(a.Pt + i * K + j) : __ldg(a.Pt + j * K + i)) : 0.f;
        const float p1 = (act && i + 1 < K) ? (fwd ? __ldg(a.Pt + (i + 1) * K + j) : __ldg(a.Pt + j * K + i + 1)) : 0.f;
        col2[i / 2] = pack2(p0, p1);
      }
      // broadcast slots [parity][warp][32 lanes]: each lane owns one word, a group reads its KP words
      // (the groups of one warp touch disjoint banks)
      float* w0 = bcS + wp * 32 + lane;
      float* w1 = w0 + 64;
      const float* r0 = bcS + wp * 32 + (lane / KP) * KP;
      const float* r1 = r0 + 64;
      const int jj = st ? j : 0;
      const int dt = fwd ? KS : -KS;
      const int tb = fwd ? 0 : T - 1;
      const float* bp = bS + (size_t)tb * KS + jj;
      float* op = (fwd ? aS : cS) + (size_t)tb * KS + jj;
      int* ep = ES + tb;
      const int de = fwd ? 1 : 0;                       // only the forward group's exponents are kept
      float v, pend;
      if (fwd) { v = act ? __ldg(a.pi0 + j) * bp[0] : 0.f; pend = v; }
      else { v = act ? bp[0] : 0.f; pend = act ? 1.f : 0.f; }
      *w1 = v;                                          // step s reads parity s & 1
      int xa = FUSED_XTB, da = 0, E = 0;
      int s = 1;
      float bn = (T > 1 && st) ? bp[dt] : 0.f;
      // each step first flushes the previous step's table entry / exponent (op, ep point at them)
      for (; s + 3 < T; s += 4) {                       // s is odd here: parities 1,0,1,0
        const float b0 = bn;
        const float b1 = st ? bp[2 * dt] : 0.f, b2 = st ? bp[3 * dt] : 0.f, b3 = st ? bp[4 * dt] : 0.f;
        bn = (s + 4 < T && st) ? bp[5 * dt] : 0.f;
        chain_step<KP>(v, col2, b0, fwd, st, r1, w0, op, ep, lead, pend, xa, da, E);
        chain_step<KP>(v, col2, b1, fwd, st, r0, w1, op + dt, ep + de, lead, pend, xa, da, E);
        chain_step<KP>(v, col2, b2, fwd, st, r1, w0, op + 2 * dt, ep + 2 * de, lead, pend, xa, da, E);
        chain_step<KP>(v, col2, b3, fwd, st, r0, w1, op + 3 * dt, ep + 3 * de, lead, pend, xa, da, E);
        bp += 4 * dt; op += 4 * dt; ep += 4 * de;
      }
      for (; s < T; ++s) {
        const float b0 = bn;
        bn = (s + 1 < T && st) ? bp[2 * dt] : 0.f;
        if (s & 1) chain_step<KP>(v, col2, b0, fwd, st, r1, w0, op, ep, lead, pend, xa, da, E);
        else chain_step<KP>(v, col2, b0, fwd, st, r0, w1, op, ep, lead, pend, xa, da, E);
        bp += dt; op += dt; ep += de;
      }
      if (st) *op = pend;                               // entry of the last step
      if (lead) *ep = E;
    }
  }
  __syncthreads();
  FUSED_STAMP(2);

  // ---------------------------------------------------------------- phase C1: marginals
  // KP/4 lanes per row, one float4 of alpha~ and beta~ each: a warp touches contiguous shared memory
  {
    constexpr int LPR = KP / 4, RPW = 32 / LPR;
    const int lane = tid & 31, wp = tid >> 5;
    const int sub = lane % LPR, rsub = lane / LPR;
    for (int row0 = wp * RPW; row0 < T; row0 += (FUSED_NT / 32) * RPW) {     // warp-uniform trip count
      const int row = row0 + rsub;
      const bool ok = row < T && 4 * sub < KS;
      float4 al = make_float4(0.f, 0.f, 0.f, 0.f), be = al;
      if (ok) {
        al = *reinterpret_cast<const float4*>(aS + (size_t)row * KS + 4 * sub);
        be = *reinterpret_cast<const float4*>(cS + (size_t)row * KS + 4 * sub);
      }
      float4 p = make_float4(al.x * be.x, al.y * be.y, al.z * be.z, al.w * be.w);
      float sa = (al.x + al.y) + (al.z + al.w), sp = (p.x + p.y) + (p.z + p.w);
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) {
        sa += __shfl_xor_sync(0xffffffffu, sa, o);
        sp += __shfl_xor_sync(0xffffffffu, sp, o);
      }
      const float inv = 1.f / sp;
      if (ok) {
        p.x *= inv; p.y *= inv; p.z *= inv; p.w *= inv;
        *reinterpret_cast<float4*>(aS + (size_t)row * KS + 4 * sub) = p;
      }
      if (sub == 0 && row < T) ltS[row] = (double)logf(sa) + (double)ES[row] * M_LN2;
    }
  }
  __syncthreads();
  // posterior marginals out: coalesced copy of the q table
  if (a.var_x_out) {
    float* dst = a.var_x_out + (size_t)w * T * K;
    if (KS == K && (((uintptr_t)dst) & 15) == 0) {
      const float4* s4 = reinterpret_cast<const float4*>(aS);
      float4* d4 = reinterpret_cast<float4*>(dst);
      for (int i = tid; i < T * K / 4; i += FUSED_NT) d4[i] = s4[i];
    } else {
      for (int i = tid; i < T * K; i += FUSED_NT) { const int t = i / K; dst[i] = aS[(size_t)t * KS + (i - t * K)]; }
    }
  }
  // log normalisers: logZ = lt[T-1] + sum_t mx[t];  Q4 = sum_t (lt[t] + sum_{s<=t} mx[s])
  if (tid < 32) {
    double smx = 0.0, q4 = 0.0;
    for (int t = tid; t < T; t += 32) { smx += mxS[t]; q4 += ltS[t] + (double)(T - t) * mxS[t]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      smx += __shfl_xor_sync(0xffffffffu, smx, o);
      q4 += __shfl_xor_sync(0xffffffffu, q4, o);
    }
    if (tid == 0) {
      const double lz = ltS[T - 1] + smx;
      a.seq[2 * (size_t)w] = lz; a.seq[2 * (size_t)w + 1] = q4;
      atomicAdd(a.stats_out + a.o_tail, lz);
      atomicAdd(a.stats_out + a.o_tail + 1, q4);
      atomicAdd(a.stats_out + a.o_tail + 2, 1.0);
    }
  }
  if (tid >= 32 && tid < 32 + K) atomicAdd(a.stats_out + a.o_q0 + (tid - 32), (double)aS[tid - 32]);
  FUSED_STAMP(3);

  // ---------------------------------------------------------------- phase C2: transition statistic
  // A[i][j] = sum_t q[t-1][i] q[t][j]: 4x4 register tiles, T split over row chunks, float32 partials
  // (<= T/chunks terms each) reduced across chunks in float64.
  const int nbi = KS / 4;
  {
    const int nb = nbi * nbi;
    int chunks = FUSED_NT / nb;
    if (chunks > (T + 7) / 8) chunks = (T + 7) / 8;
    if (chunks < 1) chunks = 1;
    const int len = (T + chunks - 1) / chunks;
    __syncthreads();                                   // b/beta (under the scratch) are dead from here on
    if (tid < nb * chunks) {
      const int c = tid / nb, blk = tid - c * nb;
      const int i0 = (blk / nbi) * 4, j0 = (blk % nbi) * 4;
      float acc[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) acc[u] = 0.f;
      const int tb = c * len, te = min(T, tb + len);
      int t = tb;
      if (t == 0 && te > 0) {                            // pair (T-1, 0): the reference's wrap-around
        if (a.wrap) {
          const float4 pv = *reinterpret_cast<const float4*>(aS + (size_t)(T - 1) * KS + i0);
          const float4 cv = *reinterpret_cast<const float4*>(aS + j0);
          const float pa[4] = {pv.x, pv.y, pv.z, pv.w}, ca[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v2 = 0; v2 < 4; ++v2) acc[u * 4 + v2] = pa[u] * ca[v2];
        }
        t = 1;
      }
      const float* pp = aS + (size_t)(t - 1) * KS + i0;
      const float* pc = aS + (size_t)t * KS + j0;
#pragma unroll 4
      for (; t < te; ++t) {
        const float4 pv = *reinterpret_cast<const float4*>(pp);
        const float4 cv = *reinterpret_cast<const float4*>(pc);
        pp += KS; pc += KS;
        const float pa[4] = {pv.x, pv.y, pv.z, pv.w}, ca[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v2 = 0; v2 < 4; ++v2) acc[u * 4 + v2] = fmaf(pa[u], ca[v2], acc[u * 4 + v2]);
      }
#pragma unroll
      for (int u = 0; u < 16; u += 4)
        *reinterpret_cast<float4*>(red + (size_t)tid * 16 + u) = make_float4(acc[u], acc[u + 1], acc[u + 2], acc[u + 3]);
    }
    __syncthreads();
    for (int e = tid; e < nb * 16; e += FUSED_NT) {
      const int blk = e / 16, u = e - blk * 16;
      const int i = (blk / nbi) * 4 + u / 4, jq = (blk % nbi) * 4 + (u & 3);
      if (i < K && jq < K) {
        double tot = 0.0;
        for (int c = 0; c < chunks; ++c) tot += (double)red[((size_t)c * nb + blk) * 16 + u];
        if (a.add_prior) tot += a.prior_tran[i * K + jq] - 1.0;
        atomicAdd(a.stats_out + i * K + jq, tot);
      }
    }
  }
  FUSED_STAMP(4);

  // ---------------------------------------------------------------- phase C3: emission statistics
  // columns of the float32 window tile: [x_0..x_{D-1} (zero on dropped rows) | w | 0-pad]
  //   S1[k][c] = sum_t q[t][k] X[t][c]                      -> sx (c < D), n (c = D)
  //   diagonal: S2[k][d] = sum_t q[t][k] X[t][d]^2          -> sxx   (same sweep as S1)
  //   full:     S2[k][d][e] = sum_t q[t][k] X[t][d] X[t][e] -> sxx   (one extra sweep per d)
  {
    {                                                   // refill the window (L2-resident by now) as float32
      const int64_t e0 = s0 * D;
      const int n = T * D;
      bool vec = false;
      if (a.dtype == SVIHMM_F32 && (D & 3) == 0) {
        const float* src = (const float*)a.obs + e0;
        if ((((uintptr_t)src) & 15) == 0) {
          vec = true;
          const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll 4
          for (int i = tid; i < n / 4; i += FUSED_NT) {
            float4 q = __ldg(s4 + i);
            const int r = (4 * i) / D, d = 4 * i - r * D;
            if (fl[r] & 1) q = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(xf + (size_t)r * DS + d) = q;
          }
        }
      }
      if (!vec) {
#pragma unroll 4
        for (int i = tid; i < n; i += FUSED_NT) {
          const int r = i / D, d = i - r * D;
          const float v = (float)ld_obs(a.obs, a.dtype, e0 + i);
          xf[(size_t)r * DS + d] = (fl[r] & 1) ? 0.f : v;
        }
      }
      const int np = DS - D;                            // w column and padding
      for (int i = tid; i < T * np; i += FUSED_NT) {
        const int r = i / np, c = D + (i - r * np);
        xf[(size_t)r * DS + c] = (c == D && !(fl[r] & 1)) ? 1.f : 0.f;
      }
    }
    const int ncb = DS / 4;
    const int nb = nbi * ncb;                           // 4x4 tiles of the (KS x DS) output
    int chunks = FUSED_NT / nb;
    if (chunks > (T + 7) / 8) chunks = (T + 7) / 8;
    if (chunks < 1) chunks = 1;
    const int len = (T + chunks - 1) / chunks;
    const bool via_red = nb * chunks <= FUSED_NT;       // else (large K*D) partials go straight to atomics
    const int npass = a.diag ? 1 : 1 + D;
    for (int pass = 0; pass < npass; ++pass) {
      __syncthreads();
      const int dsel = pass - 1;                        // full covariance: the fixed left factor x_d
      const bool sq = a.diag != 0;                      // diagonal: second accumulator set for x^2
      for (int it0 = 0; it0 < nb * chunks; it0 += FUSED_NT) {
        const int item = it0 + tid;
        if (item < nb * chunks) {
          const int c = item / nb, blk = item - c * nb;
          const int k0 = (blk / ncb) * 4, c0 = (blk % ncb) * 4;
          float acc[16], acc2[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) { acc[u] = 0.f; acc2[u] = 0.f; }
          const int tb = c * len, te = min(T, tb + len);
          const float* qp = aS + (size_t)tb * KS + k0;
          const float* xp = xf + (size_t)tb * DS + c0;
          const float* xdp = xf + (size_t)tb * DS + (dsel > 0 ? dsel : 0);
#pragma unroll 2
          for (int t = tb; t < te; ++t) {
            const float4 qv = *reinterpret_cast<const float4*>(qp);
            float4 xv = *reinterpret_cast<const float4*>(xp);
            if (pass > 0) { const float xd = *xdp; xv.x *= xd; xv.y *= xd; xv.z *= xd; xv.w *= xd; }
            qp += KS; xp += DS; xdp += DS;
            const float qa[4] = {qv.x, qv.y, qv.z, qv.w}, xa4[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
              for (int v2 = 0; v2 < 4; ++v2) acc[u * 4 + v2] = fmaf(qa[u], xa4[v2], acc[u * 4 + v2]);
            if (sq) {
              const float x2[4] = {xv.x * xv.x, xv.y * xv.y, xv.z * xv.z, xv.w * xv.w};
#pragma unroll
              for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v2 = 0; v2 < 4; ++v2) acc2[u * 4 + v2] = fmaf(qa[u], x2[v2], acc2[u * 4 + v2]);
            }
          }
          if (via_red) {
#pragma unroll
            for (int u = 0; u < 16; u += 4) {
              *reinterpret_cast<float4*>(red + (size_t)item * 32 + u) = make_float4(acc[u], acc[u + 1], acc[u + 2], acc[u + 3]);
              *reinterpret_cast<float4*>(red + (size_t)item * 32 + 16 + u) = make_float4(acc2[u], acc2[u + 1], acc2[u + 2], acc2[u + 3]);
            }
          } else {
#pragma unroll
            for (int u = 0; u < 16; ++u) {
              const int kq = k0 + u / 4, cc = c0 + (u & 3);
              if (kq < K) {
                if (pass == 0) {
                  if (cc < D) atomicAdd(a.stats_out + a.o_sx + (size_t)kq * D + cc, (double)acc[u]);
                  else if (cc == D) atomicAdd(a.stats_out + a.o_n + kq, (double)acc[u]);
                  if (sq && cc < D) atomicAdd(a.stats_out + a.o_sxx + (size_t)kq * D + cc, (double)acc2[u]);
                } else if (cc < D) {
                  atomicAdd(a.stats_out + a.o_sxx + ((size_t)kq * D + dsel) * D + cc, (double)acc[u]);
                }
              }
            }
          }
        }
      }
      if (via_red) {
        __syncthreads();
        for (int e = tid; e < nb * 32; e += FUSED_NT) {
          const int blk = e / 32, u2 = e - blk * 32, second = u2 >> 4, u = u2 & 15;
          const int kq = (blk / ncb) * 4 + u / 4, cc = (blk % ncb) * 4 + (u & 3);
          if (kq < K && cc <= D && (!second || (sq && cc < D))) {
            double tot = 0.0;
            for (int c = 0; c < chunks; ++c) tot += (double)red[((size_t)c * nb + blk) * 32 + u2];
            if (second) atomicAdd(a.stats_out + a.o_sxx + (size_t)kq * D + cc, tot);
            else if (pass == 0) {
              if (cc < D) atomicAdd(a.stats_out + a.o_sx + (size_t)kq * D + cc, tot);
              else atomicAdd(a.stats_out + a.o_n + kq, tot);
            } else if (cc < D) {
              atomicAdd(a.stats_out + a.o_sxx + ((size_t)kq * D + dsel) * D + cc, tot);
            }
          }
        }
      }
    }
  }
  FUSED_STAMP(5);
}
