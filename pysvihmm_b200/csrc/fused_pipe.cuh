// Pipelined single-kernel E-step for K <= 32 (supersedes the phase-by-phase k_estep_fused): the
// same arithmetic, but the three phases of a window overlap inside its CTA.
//
//   warps 1..7 (224 "workers")   phase A: expected log-likelihoods -> b table, rows taken OUTSIDE-IN
//                                (row p and row T-1-p per pair index p) straight from global memory,
//                                one mbarrier arrival per round of 112 pairs
//   warp 0 (the "chain" warp)    phase B: forward chain in lanes 0..15, backward chain in lanes 16..31
//                                (fused.cuh: smem broadcast + packed FFMA2 + power-of-two rescaling);
//                                starts as soon as round 0 of phase A has landed and signals an
//                                mbarrier every FP_TB steps once the two chains have crossed
//   warps 1..7 again             phase C: rows between the two chain heads are final; for every
//                                signalled tile the workers form the marginals INSIDE-OUT, write them
//                                out, and accumulate the transition / emission statistics in
//                                registers (4x4 tiles); only the cross-thread reduction and the
//                                float64 atomics remain after the chain has finished.
// The kernel time is therefore ~ (first phase-A round) + chain + (last tile + reduction) instead of
// A + B + C.  No observation staging in shared memory: phase A and C read the window rows from
// global memory (HBM once, then L2), so HBM traffic stays at the algorithmic T*D + T*K.
#pragma once
#include "fused.cuh"

#define FP_NT 256
#define FP_NW 224            // worker threads when one warp runs both chains (K <= 16); 192 for 16 < K <= 32
#define FP_TB 64             // chain steps per hand-off tile
#define FP_LA 8              // b-table lookahead of the chain (steps)
#define FP_MAXBAR 48

struct PipeSmem {
  int KS;
  size_t b, c, a, red, ES, par, flags, bc, bar, misc, dum, total;
};

__host__ __device__ inline PipeSmem pipe_smem_layout(int T, int K, int D, int tri, int diag) {
  PipeSmem s;
  s.KS = (K + 3) & ~3;
  const size_t R = (size_t)T * s.KS * sizeof(float);
  const size_t redb = (size_t)FP_NW * 48 * sizeof(float);
  s.b = 0; s.c = R; s.red = 0;                       // the reduction buffer overlays [b | beta] at the end
  s.a = 2 * R > redb ? 2 * R : fused_al16(redb);
  s.ES = fused_al16(s.a + R);
  const size_t params = diag ? ((size_t)s.KS * D * 16 + (size_t)s.KS * 8) : ((size_t)K * (tri + D + 1) * 8);
  s.par = fused_al16(s.ES + (size_t)T * 4);
  s.flags = fused_al16(s.par + params);
  s.bc = fused_al16(s.flags + (size_t)T);
  s.bar = s.bc + 2 * 2 * 32 * sizeof(float);
  s.misc = s.bar + FP_MAXBAR * 8;                    // 8 doubles of scalars
  s.dum = s.misc + 8 * 8;                            // 64 floats + 2 ints: sink of suppressed stores
  s.total = s.dum + 66 * 4;
  s.total = fused_al16(s.total);
  return s;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, int parity) {
  unsigned ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bar_workers(const int n) { asm volatile("bar.sync 1, %0;" ::"r"(n) : "memory"); }

// one float4 of the phase-C feature row of window row t: columns [c0, c0+4) of [x_0..x_{D-1} | w | 0..]
__device__ __forceinline__ float4 pipe_xrow4(const FusedArgs& a, const int64_t s0, const int t, const int c0,
                                             const bool drop, const bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (drop) return v;
  const int D = a.D;
  if (vec && c0 < D) {
    // asm volatile: the compiler must not sink this load to its first use (it is issued one tile ahead)
    const float* p = (const float*)a.obs + (s0 + t) * D + c0;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
  }
  float e[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int c = c0 + u;
    e[u] = c < D ? (float)ld_obs(a.obs, a.dtype, (s0 + t) * D + c) : (c == D ? 1.f : 0.f);
  }
  return make_float4(e[0], e[1], e[2], e[3]);
}

template <int KP>
__global__ void __launch_bounds__(FP_NT, 2) k_estep_pipe(const FusedArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int T = a.T, K = a.K, D = a.D;
  const PipeSmem L = pipe_smem_layout(T, K, D, a.tri, a.diag);
  const int KS = L.KS;
  float* bS = reinterpret_cast<float*>(smem + L.b);
  float* cS = reinterpret_cast<float*>(smem + L.c);
  float* aS = reinterpret_cast<float*>(smem + L.a);
  float* red = reinterpret_cast<float*>(smem + L.red);
  int* ES = reinterpret_cast<int*>(smem + L.ES);
  double* parS = reinterpret_cast<double*>(smem + L.par);
  unsigned char* fl = smem + L.flags;
  float* bcS = reinterpret_cast<float*>(smem + L.bc);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + L.bar);
  double* misc = reinterpret_cast<double*>(smem + L.misc);     // [0] lt[T-1], [1..3] sums
  float* dumf = reinterpret_cast<float*>(smem + L.dum);
  int* dumi = reinterpret_cast<int*>(smem + L.dum + 64 * 4);
  const int tid = threadIdx.x, w = blockIdx.x, lane = tid & 31, wp = tid >> 5;
  const int64_t s0 = a.starts[w];
#define PIPE_STAMP(i, who) do { if (a.dbg && tid == (who)) a.dbg[(size_t)w * 16 + (i)] = clock64(); } while (0)
  PIPE_STAMP(0, 0);
  constexpr int NCW = KP == 32 ? 2 : 1;                        // chain warps (K > 16: one warp per direction)
  // Workers are the warps that do NOT share a scheduler (warp id mod 4) with a chain warp: the chain
  // issues ~40 dependent instructions per step, and a busy warp on the same scheduler delays each of
  // them (measured 110 -> 145 cycles per step).  The remaining warp(s) only help in the final reduction.
  constexpr int NWW = 2 * (4 - NCW);                           // worker warps: {1,2,3,5,6,7} or {2,3,6,7}
  constexpr int NWK = 32 * NWW;                                // worker threads
  constexpr int FP_PR = NWK / 2;                               // pair indices per phase-A round

  // geometry of the hand-offs
  const int npairs = (T + 1) / 2;                               // pair index p <-> rows p and T-1-p
  const int nroundsA = (npairs + FP_PR - 1) / FP_PR;
  const int h = T / 2;                                          // first step after which rows are final: ceil((T-1)/2)
  const int ntiles = (T - h + FP_TB - 1) / FP_TB;               // steps h..T-1 in tiles of FP_TB (T = 1: one tile)
  unsigned long long* barA = bars;                              // [nroundsA]
  unsigned long long* barB = bars + nroundsA;                   // [ntiles]

  // touch this thread's first phase-A row early so that the HBM latency overlaps the setup below
  if (tid >= 32 * NCW) {
    const int wt0 = tid - 32 * NCW;
    const int row0 = wt0 < T ? wt0 : 0;
    asm volatile("prefetch.global.L2 [%0];" ::"l"((const char*)a.obs + (size_t)(s0 + row0) * D * (a.dtype == SVIHMM_F32 ? 4 : 8)));
    const int row1 = T - 1 - row0 > 0 ? T - 1 - row0 : 0;
    asm volatile("prefetch.global.L2 [%0];" ::"l"((const char*)a.obs + (size_t)(s0 + row1) * D * (a.dtype == SVIHMM_F32 ? 4 : 8)));
  }
  // ---- setup: emission constants, barriers -----------------------------------------------------
  if (a.diag) {
    for (int i = tid; i < KS * D; i += FP_NT) {                 // [d][KS], zero-padded columns
      const int d = i / KS, kk = i - d * KS;
      const bool in = kk < K;
      parS[2 * i] = in ? a.Rs[2 * (d * K + kk)] : 0.0;
      parS[2 * i + 1] = in ? a.Rs[2 * (d * K + kk) + 1] : 0.0;
    }
    for (int kk = tid; kk < KS; kk += FP_NT) parS[2 * KS * D + kk] = kk < K ? a.ck[kk] : 0.0;
  } else {
    const int np = a.tri + D + 1;
    for (int i = tid; i < K * np; i += FP_NT) {
      const int kk = i / np, p = i - kk * np;
      parS[i] = p < a.tri ? a.Rs[(size_t)kk * a.tri + p] : (p < a.tri + D ? a.gk[(size_t)kk * D + (p - a.tri)] : a.ck[kk]);
    }
  }
  if (tid < nroundsA) mbar_init(barA + tid, NWK);
  if (tid >= 32 && tid < 32 + ntiles) mbar_init(barB + (tid - 32), 32 * NCW);
  if (tid >= 64 && tid < 72) misc[tid - 64] = 0.0;
  __syncthreads();

  if (wp < NCW) {
    // =============================================================== chain warp(s) (phase B)
    const int j = lane % KP, grp = KP == 32 ? wp : lane / KP;
    const bool live = grp < 2;                                   // lane groups >= 2 (KP < 16) idle along
    const bool fw = grp == 0;
    const bool ac = live && j < K;
    const bool sv = live && j < KS;                              // columns K..KS-1 of the tables are kept at zero
    const bool ld = fw && j == 0;
    unsigned long long col2[KP / 2];
#pragma unroll
    for (int i = 0; i < KP; i += 2) {
      const float p0 = (ac && i < K) ? (fw ? __ldg(a.Pt + i * K + j) : __ldg(a.Pt + j * K + i)) : 0.f;
      const float p1 = (ac && i + 1 < K) ? (fw ? __ldg(a.Pt + (i + 1) * K + j) : __ldg(a.Pt + j * K + i + 1)) : 0.f;
      col2[i / 2] = pack2(p0, p1);
    }
    // broadcast slots [parity][chain warp][32 lanes]
    float* w0 = bcS + wp * 32 + lane; float* w1 = w0 + 64;
    const float* r0 = bcS + wp * 32 + (lane / KP) * KP; const float* r1 = r0 + 64;
    const int jj = sv ? j : 0;
    const int dt = fw ? KS : -KS;
    const int tb = fw ? 0 : T - 1;
    const float* bp = bS + (size_t)tb * KS + jj;
    float* op = (fw ? aS : cS) + (size_t)tb * KS + jj;
    int* ep = ES + tb;
    const int de = fw ? 1 : 0;                                   // only the forward chain's exponents are kept
    const float pi0j = (fw && ac) ? __ldg(a.pi0 + j) : 0.f;
    mbar_wait(barA, 0);                                          // round 0 of phase A: rows 0.. and T-1..
    PIPE_STAMP(1, 0);
    float v, pend;
    if (fw) { v = ac ? pi0j * bp[0] : 0.f; pend = v; }
    else { v = ac ? bp[0] : 0.f; pend = ac ? 1.f : 0.f; }
    *w1 = v;                                                     // step s reads parity s & 1
    int xa = FUSED_XTB, da = 0, E = 0;
    int s = 1, roundA = 1, tile = 0;
    bool skip = false;      // the entry of the previous step was flushed at a hand-off: do not store it again
                            // (a worker may already have turned that row into marginals in place)
    auto tile_end = [&](int i) { const int e = h - 1 + FP_TB * (i + 1); return e < T - 1 ? e : T - 1; };
    if (tile_end(0) <= 0) {                                      // T == 1: the only row is final at once
      if (sv) *op = pend;
      if (ld) *ep = E;
      __syncwarp();
      mbar_arrive(barB);
      tile = 1;
    }
    while (s < T) {
      if (roundA < nroundsA && s + FP_LA >= FP_PR * roundA) { mbar_wait(barA + roundA, 0); ++roundA; }
      int seg_end = T - 1;
      if (roundA < nroundsA) seg_end = min(seg_end, FP_PR * roundA - FP_LA - 1);
      if (tile < ntiles) seg_end = min(seg_end, tile_end(tile));
      if (seg_end < s) seg_end = s;
      // steps s .. seg_end
      float bn = sv ? bp[dt] : 0.f;
      const bool odd = s & 1;
      const float* rA = odd ? r1 : r0; float* wA = odd ? w0 : w1;   // step s reads parity s & 1, writes the other
      const float* rB = odd ? r0 : r1; float* wB = odd ? w1 : w0;
      float* opf = skip ? dumf + wp * 32 + lane : op;
      int* epf = skip ? dumi : ep;
      skip = false;
      for (; s + 3 <= seg_end; s += 4) {
        const float b0 = bn;
        const float b1 = sv ? bp[2 * dt] : 0.f, b2 = sv ? bp[3 * dt] : 0.f, b3 = sv ? bp[4 * dt] : 0.f;
        bn = (s + 4 <= seg_end && sv) ? bp[5 * dt] : 0.f;
        chain_step<KP>(v, col2, b0, fw, sv, rA, wA, opf, epf, ld, pend, xa, da, E);
        chain_step<KP>(v, col2, b1, fw, sv, rB, wB, op + dt, ep + de, ld, pend, xa, da, E);
        chain_step<KP>(v, col2, b2, fw, sv, rA, wA, op + 2 * dt, ep + 2 * de, ld, pend, xa, da, E);
        chain_step<KP>(v, col2, b3, fw, sv, rB, wB, op + 3 * dt, ep + 3 * de, ld, pend, xa, da, E);
        bp += 4 * dt; op += 4 * dt; ep += 4 * de;
        opf = op; epf = ep;
      }
      for (; s <= seg_end; ++s) {
        const float b0 = bn;
        bn = (s + 1 <= seg_end && sv) ? bp[2 * dt] : 0.f;
        if (s & 1) chain_step<KP>(v, col2, b0, fw, sv, r1, w0, opf, epf, ld, pend, xa, da, E);
        else chain_step<KP>(v, col2, b0, fw, sv, r0, w1, opf, epf, ld, pend, xa, da, E);
        bp += dt; op += dt; ep += de;
        opf = op; epf = ep;
      }
      if (tile < ntiles && seg_end == tile_end(tile)) {
        if (sv) *op = pend;                                      // flush the entry of step seg_end
        if (ld) *ep = E;
        skip = true;
        __syncwarp();
        mbar_arrive(barB + tile);
        ++tile;
      }
    }
    // (the last tile ends at step T-1, so every entry has been flushed by a hand-off)
    PIPE_STAMP(2, 0);
  } else if ((wp & 3) >= NCW) {
    // =============================================================== workers
    const int wwp = (wp >> 2) * (4 - NCW) + (wp & 3) - NCW;       // rank among the worker warps
    const int wt = wwp * 32 + lane;
    // ---------------------------------------------------------------- phase A, outside-in
    double smx = 0.0, smxT = 0.0;                                // sum mx[t], sum (T - t) mx[t] over this thread's rows
    {
      const bool vecx = a.dtype == SVIHMM_F32 && (D & 3) == 0 && ((((uintptr_t)a.obs) & 15) == 0);
      for (int r = 0; r < nroundsA; ++r) {
        const int side = wt / FP_PR, p = r * FP_PR + (wt - side * FP_PR);
        const int row = side == 0 ? p : T - 1 - p;
        const bool valid = p < npairs && (side == 0 || row > p);
        if (valid) {
          const int64_t e0 = (s0 + row) * D;
          bool bad = false;
          double ll[KP];
          if (a.diag) {
#pragma unroll
            for (int k = 0; k < KP; ++k) ll[k] = k < KS ? parS[2 * KS * D + k] : 0.0;
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
                  const double2* pp = reinterpret_cast<const double2*>(parS) + (size_t)(d0 + u) * KS;
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
              }
            }
          } else {
            for (int d = 0; d < D; ++d) bad |= isnan(ld_obs(a.obs, a.dtype, e0 + d));
            const int np = a.tri + D + 1;
#pragma unroll
            for (int k = 0; k < KP; ++k) {
              ll[k] = -INFINITY;
              if (k < K) {
                const double* pr = parS + (size_t)k * np;
                double acc = 0.0;
                int o = 0;
                for (int i = 0; i < D; ++i) {
                  double sacc = -pr[a.tri + i];
                  for (int jx = 0; jx <= i; ++jx) sacc = fma(pr[o + jx], ld_obs(a.obs, a.dtype, e0 + jx), sacc);
                  o += i + 1;
                  acc = fma(sacc, sacc, acc);
                }
                ll[k] = pr[a.tri + D] - acc;
              }
            }
          }
          const bool mk = a.mask && a.mask[s0 + row];
          const bool noev = bad || (a.mask_ll && mk);
          fl[row] = (unsigned char)(((bad || mk) ? 1 : 0) | (noev ? 2 : 0));
          double m = -INFINITY;
#pragma unroll
          for (int k = 0; k < KP; ++k) { if (k < K) { if (noev) ll[k] = 0.0; m = fmax(m, ll[k]); } }
          smx += m; smxT += (double)(T - row) * m;
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
        mbar_arrive(barA + r);
      }
    }

    PIPE_STAMP(3, 32 * (NCW == 1 ? 1 : 2));
    // ---------------------------------------------------------------- phase C, inside-out
    const int nbi = KS / 4;
    const int DS = (D + 1 + 3) & ~3, ncb = DS / 4;
    const int nbA = nbi * nbi, slotsA = max(1, NWK / nbA);
    const int nbE = nbi * ncb, slotsE = max(1, NWK / nbE);
    const int nb2 = a.diag ? nbE : nbi * D * ncb, slots2 = max(1, NWK / nb2);
    const bool roleA = wt < nbA * slotsA, roleE = wt < nbE * slotsE, role2 = !a.diag && wt < nb2 * slots2;
    const int blkA = wt % nbA, slotA = wt / nbA, i0 = (blkA / nbi) * 4, j0 = (blkA % nbi) * 4;
    const int blkE = wt % nbE, slotE = wt / nbE, k0E = (blkE / ncb) * 4, c0E = (blkE % ncb) * 4;
    const int blk2 = wt % nb2, slot2 = wt / nb2;
    const int k02 = (blk2 / (D * ncb)) * 4, d2 = (blk2 / ncb) % D, e02 = (blk2 % ncb) * 4;
    const bool vecx = a.dtype == SVIHMM_F32 && (D & 3) == 0 && ((((uintptr_t)a.obs) & 15) == 0);
    float accA[16], accE[16], acc2[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) { accA[u] = 0.f; accE[u] = 0.f; acc2[u] = 0.f; }
    double ltsum = 0.0;
    constexpr int LPR = KP / 4, RPW = 32 / LPR;
    const int sub = lane % LPR, rsub = lane / LPR;
    int lo = 0, hi = -1;
    bool first = true;
    // the window rows a tile needs do not depend on the chain: each emission thread fetches its (up to
    // four) rows of the NEXT tile while it waits for the chain to hand that tile over
    float4 xpre[4];
    auto tile_rows = [&](int tile_, int lo_, int hi_, bool first_, int& nlo_, int& nhi_, int& plo_, int& phi_) {
      int e_ = h - 1 + FP_TB * (tile_ + 1); if (e_ > T - 1) e_ = T - 1;
      nlo_ = T - 1 - e_; nhi_ = e_;
      if (first_) { plo_ = (nlo_ + nhi_ + 1) / 2; phi_ = plo_ - 1; } else { plo_ = lo_; phi_ = hi_; }
    };
    auto prefetch_tile = [&](int tile_, int lo_, int hi_, bool first_) {
      int nlo_, nhi_, plo_, phi_;
      tile_rows(tile_, lo_, hi_, first_, nlo_, nhi_, plo_, phi_);
      const int nl_ = plo_ - nlo_, nnew_ = nl_ + (nhi_ - phi_);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int n = slotE + i * slotsE;
        xpre[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (roleE && n < nnew_) {
          const int row = n < nl_ ? nlo_ + n : phi_ + 1 + (n - nl_);
          xpre[i] = pipe_xrow4(a, s0, row, c0E, false, vecx);
        }
      }
    };
    prefetch_tile(0, 0, -1, true);
    long long tw = 0, t1c = 0, tbar = 0, t2c = 0, t3c = 0;
    for (int tile = 0; tile < ntiles; ++tile) {
      long long c0 = clock64();
      mbar_wait(barB + tile, 0);
      long long c1 = clock64(); tw += c1 - c0;
      int nlo, nhi, plo, phi;
      tile_rows(tile, lo, hi, first, nlo, nhi, plo, phi);
      lo = plo; hi = phi;
      const int nl = lo - nlo, nr = nhi - hi, nnew = nl + nr;
      // ---- C1: marginals of the new rows
      for (int n0 = wwp * RPW; n0 < nnew; n0 += NWW * RPW) {     // warp-uniform trip count
        const int n = n0 + rsub;
        const bool rok = n < nnew;
        const int row = n < nl ? nlo + n : hi + 1 + (n - nl);
        const bool ok = rok && 4 * sub < KS;
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
          if (a.var_x_out) {
            float* dst = a.var_x_out + ((size_t)w * T + row) * K + 4 * sub;
            if (KS == K && ((((uintptr_t)a.var_x_out) & 15) == 0)) *reinterpret_cast<float4*>(dst) = p;
            else {
              const float pe[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
              for (int u = 0; u < 4; ++u) if (4 * sub + u < K) dst[u] = pe[u];
            }
          }
        }
        if (sub == 0 && rok) {
          const double lt = (double)logf(sa) + (double)ES[row] * M_LN2;
          ltsum += lt;
          if (row == T - 1) misc[0] = lt;
        }
      }
      long long c2 = clock64(); t1c += c2 - c1;
      bar_workers(NWK);
      long long c3 = clock64(); tbar += c3 - c2;
      // ---- C2: transition pairs (u-1, u) that became available
      if (roleA) {
        const int npl = nl;                                       // u in [nlo+1, lo]
        const int ru0 = max(hi + 1, lo + 1), npr = nhi - ru0 + 1; // u in [ru0, nhi]
        const int np_ = npl + (npr > 0 ? npr : 0);
        for (int n = slotA; n < np_; n += slotsA) {
          const int u = n < npl ? nlo + 1 + n : ru0 + (n - npl);
          const float4 pv = *reinterpret_cast<const float4*>(aS + (size_t)(u - 1) * KS + i0);
          const float4 cv = *reinterpret_cast<const float4*>(aS + (size_t)u * KS + j0);
          const float pa[4] = {pv.x, pv.y, pv.z, pv.w}, ca[4] = {cv.x, cv.y, cv.z, cv.w};
#pragma unroll
          for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) accA[x * 4 + y] = fmaf(pa[x], ca[y], accA[x * 4 + y]);
        }
      }
      long long c4 = clock64(); t2c += c4 - c3;
      // ---- C3: emission statistics of the new rows
      if (roleE) {
        int ix = 0;
        for (int n = slotE; n < nnew; n += slotsE, ++ix) {
          const int row = n < nl ? nlo + n : hi + 1 + (n - nl);
          const bool drop = fl[row] & 1;
          const float4 qv = *reinterpret_cast<const float4*>(aS + (size_t)row * KS + k0E);
          float4 xv;
          if (ix < 4) {
            xv = ix == 0 ? xpre[0] : (ix == 1 ? xpre[1] : (ix == 2 ? xpre[2] : xpre[3]));
            if (drop) xv = make_float4(0.f, 0.f, 0.f, 0.f);
          } else xv = pipe_xrow4(a, s0, row, c0E, drop, vecx);
          const float qa[4] = {qv.x, qv.y, qv.z, qv.w}, xa4[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
          for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) accE[x * 4 + y] = fmaf(qa[x], xa4[y], accE[x * 4 + y]);
          if (a.diag) {
            const float x2[4] = {xv.x * xv.x, xv.y * xv.y, xv.z * xv.z, xv.w * xv.w};
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
              for (int y = 0; y < 4; ++y) acc2[x * 4 + y] = fmaf(qa[x], x2[y], acc2[x * 4 + y]);
          }
        }
      }
      if (role2) {
        for (int n = slot2; n < nnew; n += slots2) {
          const int row = n < nl ? nlo + n : hi + 1 + (n - nl);
          const bool drop = fl[row] & 1;
          const float4 qv = *reinterpret_cast<const float4*>(aS + (size_t)row * KS + k02);
          float4 xv = pipe_xrow4(a, s0, row, e02, drop, vecx);
          const float xd = drop ? 0.f : (float)ld_obs(a.obs, a.dtype, (s0 + row) * D + d2);
          xv.x *= xd; xv.y *= xd; xv.z *= xd; xv.w *= xd;
          const float qa[4] = {qv.x, qv.y, qv.z, qv.w}, xa4[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
          for (int x = 0; x < 4; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc2[x * 4 + y] = fmaf(qa[x], xa4[y], acc2[x * 4 + y]);
        }
      }
      lo = nlo; hi = nhi; first = false;
      if (tile + 1 < ntiles) prefetch_tile(tile + 1, lo, hi, false);
      t3c += clock64() - c4;
    }
    if (a.dbg && wt == 0) { long long* dq = a.dbg + (size_t)w * 16; dq[8] = tw; dq[9] = t1c; dq[10] = tbar; dq[11] = t2c; dq[12] = t3c; }
    PIPE_STAMP(4, 32 * (NCW == 1 ? 1 : 2));
    // partial results of this thread -> shared memory (the tables under `red` are dead: the chain has
    // finished (last tile signalled) and every worker is past its last read of b / beta)
    bar_workers(NWK);
    {
      float* rp = red + (size_t)wt * 48;
#pragma unroll
      for (int u = 0; u < 16; u += 4) {
        *reinterpret_cast<float4*>(rp + u) = make_float4(accA[u], accA[u + 1], accA[u + 2], accA[u + 3]);
        *reinterpret_cast<float4*>(rp + 16 + u) = make_float4(accE[u], accE[u + 1], accE[u + 2], accE[u + 3]);
        *reinterpret_cast<float4*>(rp + 32 + u) = make_float4(acc2[u], acc2[u + 1], acc2[u + 2], acc2[u + 3]);
      }
    }
    // log-normaliser pieces: reduce (sum mx, sum (T-t) mx, sum lt) over the workers
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      smx += __shfl_xor_sync(0xffffffffu, smx, o);
      smxT += __shfl_xor_sync(0xffffffffu, smxT, o);
      ltsum += __shfl_xor_sync(0xffffffffu, ltsum, o);
    }
    if (lane == 0) { atomicAdd(misc + 1, smx); atomicAdd(misc + 2, smxT); atomicAdd(misc + 3, ltsum); }
  }
  __syncthreads();

  // ---------------------------------------------------------------- reduction + float64 atomics
  {
    const int nbi = KS / 4;
    const int DS = (D + 1 + 3) & ~3, ncb = DS / 4;
    const int nbA = nbi * nbi, slotsA = max(1, NWK / nbA);
    const int nbE = nbi * ncb, slotsE = max(1, NWK / nbE);
    const int nb2 = a.diag ? nbE : nbi * D * ncb, slots2 = max(1, NWK / nb2);
    for (int e = tid; e < nbA * 16; e += FP_NT) {
      const int blk = e / 16, u = e - blk * 16;
      const int i = (blk / nbi) * 4 + u / 4, jq = (blk % nbi) * 4 + (u & 3);
      if (i < K && jq < K) {
        double tot = 0.0;
        for (int c = 0; c < slotsA; ++c) tot += (double)red[((size_t)c * nbA + blk) * 48 + u];
        if (a.wrap) tot += (double)(aS[(size_t)(T - 1) * KS + i] * aS[jq]);      // pair (T-1, 0), quirk Q2
        if (a.add_prior) tot += a.prior_tran[i * K + jq] - 1.0;
        atomicAdd(a.stats_out + i * K + jq, tot);
      }
    }
    for (int e = tid; e < nbE * 16; e += FP_NT) {
      const int blk = e / 16, u = e - blk * 16;
      const int kq = (blk / ncb) * 4 + u / 4, cc = (blk % ncb) * 4 + (u & 3);
      if (kq < K && cc <= D) {
        double tot = 0.0, tot2 = 0.0;
        for (int c = 0; c < slotsE; ++c) {
          tot += (double)red[((size_t)c * nbE + blk) * 48 + 16 + u];
          tot2 += (double)red[((size_t)c * nbE + blk) * 48 + 32 + u];
        }
        if (cc < D) {
          atomicAdd(a.stats_out + a.o_sx + (size_t)kq * D + cc, tot);
          if (a.diag) atomicAdd(a.stats_out + a.o_sxx + (size_t)kq * D + cc, tot2);
        } else atomicAdd(a.stats_out + a.o_n + kq, tot);
      }
    }
    if (!a.diag) {
      for (int e = tid; e < nb2 * 16; e += FP_NT) {
        const int blk = e / 16, u = e - blk * 16;
        const int kq = (blk / (D * ncb)) * 4 + u / 4, dd = (blk / ncb) % D, cc = (blk % ncb) * 4 + (u & 3);
        if (kq < K && cc < D) {
          double tot = 0.0;
          for (int c = 0; c < slots2; ++c) tot += (double)red[((size_t)c * nb2 + blk) * 48 + 32 + u];
          atomicAdd(a.stats_out + a.o_sxx + ((size_t)kq * D + dd) * D + cc, tot);
        }
      }
    }
    if (tid < K) atomicAdd(a.stats_out + a.o_q0 + tid, (double)aS[tid]);
    PIPE_STAMP(5, 32);
    if (tid == 32) {
      // logZ = lt[T-1] + sum_t mx[t];  Q4 = sum_t (lt[t] + sum_{s<=t} mx[s]) = sum lt + sum (T - t) mx[t]
      const double lz = misc[0] + misc[1], q4 = misc[3] + misc[2];
      a.seq[2 * (size_t)w] = lz; a.seq[2 * (size_t)w + 1] = q4;
      atomicAdd(a.stats_out + a.o_tail, lz);
      atomicAdd(a.stats_out + a.o_tail + 1, q4);
      atomicAdd(a.stats_out + a.o_tail + 2, 1.0);
    }
  }
}
