// Pipelined single-kernel E-step for the diagonal model with K <= 32 (other models: k_estep_fused,
// the phase-by-phase version of the same arithmetic): the three phases of a window overlap inside
// its CTA.
//
//   worker warps {1,2,3,5,6,7}   phase A: expected log-likelihoods -> b table, rows taken OUTSIDE-IN
//   (the warps that do not       (row p and row T-1-p per pair index p) straight from global memory,
//   share the chain's            one mbarrier arrival per round of 96 pairs
//   scheduler)
//   warp 0 (the "chain" warp)    phase B: forward chain in lanes 0..15, backward chain in lanes 16..31
//                                (fused.cuh: smem broadcast + packed FFMA2 + power-of-two rescaling);
//                                starts as soon as round 0 of phase A has landed and signals an
//                                mbarrier every FP_TB steps once the two chains have crossed
//   worker warps again           phase C: rows between the two chain heads are final; for every
//                                signalled tile the workers form the marginals INSIDE-OUT, write them
//                                out, and accumulate the transition / emission statistics on the
//                                TENSOR CORES (mma.m16n8k8, 3xTF32, float32 accumulators in
//                                registers); only the cross-warp reduction and the float64 atomics
//                                remain after the chain has finished.
// The kernel time is therefore ~ (first phase-A round) + chain + (last tile + reduction) instead of
// A + B + C (measured at c2: chain ends at 75.6k cycles, CTA done at 86k; the scalar phase C of the
// first version finished at 103k).  No observation staging in shared memory: phase A and C read the
// window rows from global memory (HBM once, then L2), so HBM traffic stays at the algorithmic
// T*D + T*K.
#pragma once
#include "fused.cuh"

#define FP_NT 256
#define FP_NW 224            // worker threads when one warp runs both chains (K <= 16); 192 for 16 < K <= 32
#define FP_TB 48             // chain steps per hand-off tile (96 rows = two rounds of 8-row MMA groups over 6 warps; measured 64: 59.9, 48: 58.4, 32: 60.4 us)
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
// waiting for a hand-off that is microseconds away: back off between polls so that the spinning
// warps do not take issue slots and shared-memory bandwidth from the chain warp (ncu: 30 % of all
// executed instructions of the kernel were try_wait polls)
__device__ __forceinline__ void mbar_wait_relaxed(unsigned long long* bar, int parity) {
  unsigned ok;
  for (;;) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) break;
    __nanosleep(200);
  }
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

// 3xTF32 helpers: v = hi + lo with hi, lo representable in TF32 (10-bit mantissa)
__device__ __forceinline__ void split_tf32(const float v, unsigned& hi, unsigned& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
  const float r = v - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// KP: lane group of a chain (power of two >= K); NTE: 8-column tiles of the emission features
// [x | x^2 | w] = 2*ceil(D/8) + 1 (diagonal model only; other models use k_estep_fused)
template <int KP, int NTE>
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
  if (a.dbg && tid == 0) { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); a.dbg[(size_t)w * 16 + 13] = (long long)gt; }
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
  // Programmatic dependent launch (svihmm_svi_run): everything above (window start, L2 prefetch of the
  // first rows) does not depend on the global parameters and may overlap the update kernel; from here on
  // the parameters it wrote are read.  They are loaded through L2 (ld.global.cg): a line of the PREVIOUS
  // step's parameters may still sit in this SM's L1 when the CTA was scheduled before the update finished.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (a.diag) {
    for (int i = tid; i < KS * D; i += FP_NT) {                 // [d][KS], zero-padded columns
      const int d = i / KS, kk = i - d * KS;
      const bool in = kk < K;
      parS[2 * i] = in ? __ldcg(a.Rs + 2 * (d * K + kk)) : 0.0;
      parS[2 * i + 1] = in ? __ldcg(a.Rs + 2 * (d * K + kk) + 1) : 0.0;
    }
    for (int kk = tid; kk < KS; kk += FP_NT) parS[2 * KS * D + kk] = kk < K ? __ldcg(a.ck + kk) : 0.0;
  } else {
    const int np = a.tri + D + 1;
    for (int i = tid; i < K * np; i += FP_NT) {
      const int kk = i / np, p = i - kk * np;
      parS[i] = p < a.tri ? __ldcg(a.Rs + (size_t)kk * a.tri + p) : (p < a.tri + D ? __ldcg(a.gk + (size_t)kk * D + (p - a.tri)) : __ldcg(a.ck + kk));
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
      const float p0 = (ac && i < K) ? (fw ? __ldcg(a.Pt + i * K + j) : __ldcg(a.Pt + j * K + i)) : 0.f;
      const float p1 = (ac && i + 1 < K) ? (fw ? __ldcg(a.Pt + (i + 1) * K + j) : __ldcg(a.Pt + j * K + i + 1)) : 0.f;
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
    const float pi0j = (fw && ac) ? __ldcg(a.pi0 + j) : 0.f;
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
    // Statistics on the tensor cores: per group of 8 rows one warp forms, with mma.m16n8k8 (TF32
    // inputs split hi/lo = "3xTF32", float32 accumulate),
    //     accT[mt][j] += Q[prev rows]^T (16 states x 8) . Q[next rows] (8 x 8 states)      transitions
    //     accE[mt][j] += Q[rows]^T . F[rows],  F = [ x (ND8 tiles) | x^2 (ND8 tiles) | w ]    emissions
    // (ncu on the scalar version: 37k of the CTA's 118k executed warp instructions were phase C
    // index arithmetic around 16-FMA register tiles and the worker schedulers were ~80 % busy; one
    // group of 8 rows is now ~110 warp instructions instead of ~580).
    constexpr int MT = KP > 16 ? 2 : 1;                 // 16-state tiles
    constexpr int NTT = KP > 8 ? KP / 8 : 1;            // 8-state tiles of the "next" side
    const int g = lane >> 2, tig = lane & 3;
    const int ND8 = (D + 7) >> 3;
    float accT[MT][NTT][4], accE[MT][NTE][4];
#pragma unroll
    for (int m = 0; m < MT; ++m) {
#pragma unroll
      for (int j = 0; j < NTT; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) accT[m][j][c] = 0.f;
#pragma unroll
      for (int j = 0; j < NTE; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) accE[m][j][c] = 0.f;
    }
    double ltsum = 0.0;
    constexpr int LPR = KP / 4, RPW = 32 / LPR;
    const int sub = lane % LPR, rsub = lane / LPR;
    int lo = 0, hi = -1;
    bool first = true;
    auto tile_rows = [&](int tile_, int lo_, int hi_, bool first_, int& nlo_, int& nhi_, int& plo_, int& phi_) {
      int e_ = h - 1 + FP_TB * (tile_ + 1); if (e_ > T - 1) e_ = T - 1;
      nlo_ = T - 1 - e_; nhi_ = e_;
      if (first_) { plo_ = (nlo_ + nhi_ + 1) / 2; phi_ = plo_ - 1; } else { plo_ = lo_; phi_ = hi_; }
    };
    // q entries of one A fragment (states 16*mt + {g, g+8}) of table rows r0 / r1 (ok0/ok1: row exists)
    auto load_a = [&](const int mt, const int r0, const int r1, const bool ok0, const bool ok1, float (&av)[4]) {
      const int c0 = 16 * mt + g, c1 = c0 + 8;
      av[0] = (ok0 && c0 < KS) ? aS[(size_t)r0 * KS + c0] : 0.f;
      av[1] = (ok0 && c1 < KS) ? aS[(size_t)r0 * KS + c1] : 0.f;
      av[2] = (ok1 && c0 < KS) ? aS[(size_t)r1 * KS + c0] : 0.f;
      av[3] = (ok1 && c1 < KS) ? aS[(size_t)r1 * KS + c1] : 0.f;
    };
    long long tw = 0, t1c = 0, tbar = 0, t2c = 0, t3c = 0;
    for (int tile = 0; tile < ntiles; ++tile) {
      long long c0 = clock64();
      mbar_wait_relaxed(barB + tile, 0);
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
      // ---- C2: transition pairs (u-1, u) that became available, 8 pairs per warp per pass
      {
        const int npl = nl;                                       // u in [nlo+1, lo]
        const int ru0 = max(hi + 1, lo + 1), npr = nhi - ru0 + 1; // u in [ru0, nhi]
        const int np_ = npl + (npr > 0 ? npr : 0);
        for (int g0 = wwp * 8; g0 < np_; g0 += NWW * 8) {
          const int n0 = g0 + tig, n1 = n0 + 4;
          const bool ok0 = n0 < np_, ok1 = n1 < np_;
          const int u0 = n0 < npl ? nlo + 1 + n0 : ru0 + (n0 - npl);
          const int u1 = n1 < npl ? nlo + 1 + n1 : ru0 + (n1 - npl);
          unsigned bh[NTT][2], bl[NTT][2];
#pragma unroll
          for (int j = 0; j < NTT; ++j) {
            const int cn = 8 * j + g;
            const float b0 = (ok0 && cn < KS) ? aS[(size_t)u0 * KS + cn] : 0.f;
            const float b1 = (ok1 && cn < KS) ? aS[(size_t)u1 * KS + cn] : 0.f;
            split_tf32(b0, bh[j][0], bl[j][0]); split_tf32(b1, bh[j][1], bl[j][1]);
          }
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            float av[4];
            load_a(mt, u0 - 1, u1 - 1, ok0, ok1, av);
            unsigned ah[4], al_[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) split_tf32(av[c], ah[c], al_[c]);
#pragma unroll
            for (int j = 0; j < NTT; ++j) {
              mma_tf32(accT[mt][j], al_, bh[j]); mma_tf32(accT[mt][j], ah, bl[j]); mma_tf32(accT[mt][j], ah, bh[j]);
            }
          }
        }
      }
      long long c4 = clock64(); t2c += c4 - c3;
      // ---- C3: emission statistics of the new rows (masked / NaN rows dropped: w = 0)
      for (int g0 = wwp * 8; g0 < nnew; g0 += NWW * 8) {
        const int n0 = g0 + tig, n1 = n0 + 4;
        const bool ok0 = n0 < nnew, ok1 = n1 < nnew;
        const int r0 = n0 < nl ? nlo + n0 : hi + 1 + (n0 - nl);
        const int r1 = n1 < nl ? nlo + n1 : hi + 1 + (n1 - nl);
        const bool w0 = ok0 && !(fl[r0] & 1), w1 = ok1 && !(fl[r1] & 1);
        unsigned bh[NTE][2], bl[NTE][2];
#pragma unroll
        for (int j = 0; j < NTE; ++j) { bh[j][0] = bh[j][1] = bl[j][0] = bl[j][1] = 0u; }
#pragma unroll
        for (int j = 0; j < (NTE - 1) / 2; ++j) {
          if (j < ND8) {
            const int d = 8 * j + g;
            const float x0 = (w0 && d < D) ? (float)ld_obs(a.obs, a.dtype, (s0 + r0) * D + d) : 0.f;
            const float x1 = (w1 && d < D) ? (float)ld_obs(a.obs, a.dtype, (s0 + r1) * D + d) : 0.f;
            split_tf32(x0, bh[j][0], bl[j][0]); split_tf32(x1, bh[j][1], bl[j][1]);
            split_tf32(x0 * x0, bh[(NTE - 1) / 2 + j][0], bl[(NTE - 1) / 2 + j][0]);
            split_tf32(x1 * x1, bh[(NTE - 1) / 2 + j][1], bl[(NTE - 1) / 2 + j][1]);
          }
        }
        bh[NTE - 1][0] = (w0 && g == 0) ? 0x3f800000u : 0u;      // the count column: w = 1.0 (exact in TF32)
        bh[NTE - 1][1] = (w1 && g == 0) ? 0x3f800000u : 0u;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          float av[4];
          load_a(mt, r0, r1, ok0, ok1, av);
          unsigned ah[4], al_[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) split_tf32(av[c], ah[c], al_[c]);
#pragma unroll
          for (int j = 0; j < NTE - 1; ++j) {
            mma_tf32(accE[mt][j], al_, bh[j]); mma_tf32(accE[mt][j], ah, bl[j]); mma_tf32(accE[mt][j], ah, bh[j]);
          }
          mma_tf32(accE[mt][NTE - 1], al_, bh[NTE - 1]); mma_tf32(accE[mt][NTE - 1], ah, bh[NTE - 1]);
        }
      }
      lo = nlo; hi = nhi; first = false;
      t3c += clock64() - c4;
    }
    if (a.dbg && wt == 0) { long long* dq = a.dbg + (size_t)w * 16; dq[8] = tw; dq[9] = t1c; dq[10] = tbar; dq[11] = t2c; dq[12] = t3c; }
    PIPE_STAMP(4, 32 * (NCW == 1 ? 1 : 2));
    // accumulator fragments of this warp -> shared memory (the tables under `red` are dead: the chain
    // has finished (last tile signalled) and every worker is past its last read of b / beta)
    bar_workers(NWK);
    {
      float* rp = red + (size_t)wwp * (MT * (NTT + NTE) * 4) * 32 + lane;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int j = 0; j < NTT; ++j)
#pragma unroll
          for (int c = 0; c < 4; ++c) rp[(((mt * (NTT + NTE)) + j) * 4 + c) * 32] = accT[mt][j][c];
#pragma unroll
        for (int j = 0; j < NTE; ++j)
#pragma unroll
          for (int c = 0; c < 4; ++c) rp[(((mt * (NTT + NTE)) + NTT + j) * 4 + c) * 32] = accE[mt][j][c];
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
    constexpr int MT = KP > 16 ? 2 : 1, NTT = KP > 8 ? KP / 8 : 1, NJ = NTT + NTE;
    const int ND8 = (D + 7) >> 3;
    // element (state m, tile j, column n of the tile) sits in fragment register c of lane 4*(m%8)+n/2
    auto frag_sum = [&](const int m, const int j, const int n) {
      const int mt = m >> 4, mm = m & 15, c = 2 * (mm >> 3) + (n & 1), ln = 4 * (mm & 7) + (n >> 1);
      double tot = 0.0;
      for (int wv_ = 0; wv_ < NWW; ++wv_)
        tot += (double)red[((size_t)wv_ * (MT * NJ * 4) + ((mt * NJ + j) * 4 + c)) * 32 + ln];
      return tot;
    };
    {
    for (int e = tid; e < K * K; e += FP_NT) {
      const int i = e / K, jq = e - i * K;
      double tot = frag_sum(i, jq >> 3, jq & 7);
      if (a.wrap) tot += (double)(aS[(size_t)(T - 1) * KS + i] * aS[jq]);      // pair (T-1, 0), quirk Q2
      if (a.add_prior) tot += a.prior_tran[e] - 1.0;
      atomicAdd(a.stats_out + e, tot);
    }
    for (int e = tid; e < K * (2 * D + 1); e += FP_NT) {
      const int kq = e / (2 * D + 1), f = e - kq * (2 * D + 1);
      if (f < D) atomicAdd(a.stats_out + a.o_sx + (size_t)kq * D + f, frag_sum(kq, NTT + (f >> 3), f & 7));
      else if (f < 2 * D) {
        const int d = f - D;
        atomicAdd(a.stats_out + a.o_sxx + (size_t)kq * D + d, frag_sum(kq, NTT + (NTE - 1) / 2 + (d >> 3), d & 7));
      } else atomicAdd(a.stats_out + a.o_n + kq, frag_sum(kq, NTT + NTE - 1, 0));
    }
    if (tid < K) atomicAdd(a.stats_out + a.o_q0 + tid, (double)aS[tid]);
    }
    (void)ND8;
    PIPE_STAMP(5, 32);
    if (a.dbg && tid == 32) { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); a.dbg[(size_t)w * 16 + 14] = (long long)gt; }
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
