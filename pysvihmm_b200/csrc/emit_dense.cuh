// Diagonal expected log-likelihoods on tcgen05 for the opt-in bf16 dense path (BASELINE config 4: K = 256,
// D = 64, SVIHMM_BF16_DENSE - marginals agree with the float64 reference to ~1e-2 there, see dense.cuh):
//   ll[r][k] = ck'_k + sum_d (c2_kd x_rd^2 + c1_kd x_rd)      (expanded form of the fused kernels, global.cuh par2 / ckp;
//                                                              the reference has no diagonal mean-field model)
// is the GEMM [x^2 | x] (rows x 2D) . [c2 ; c1] (2D x K) with every operand split into bf16 hi + lo
// (v = hi + lo to 2^-17; products hi.hi + hi.lo + lo.hi, float32 accumulators in TENSOR MEMORY): |ll| ~ 1e2 comes out
// to ~3e-4 absolute, an order below what the bf16 messages of this path carry.  In float64 the phase was 1.8 ms
// of the 5.0 ms step (k_emit_diag_rb, FP64 pipe); here the MMAs of a 128-row tile take ~3 k cycles.
//   one persistent CTA per SM, 256 threads; tile = 128 consecutive rows of one window
//   TMA       cp.async.bulk.tensor.2d: the 128 x D block of observations (tensor map over the (T_full, D) series)
//   operands  B = [c2 | c1] of all K states, hi and lo, K-major SWIZZLE_128B, resident for the whole kernel
//             (128 KB at K = 256, D = 64); A = the tile's features [x^2 | x], hi and lo (64 KB), regenerated per tile
//   MMA       thread 0: 2D/16 k-steps x 3 terms of tcgen05.mma.kind::f16 (bf16, M = 128, N = K) into TMEM stage
//             (tile & 1); tcgen05.commit -> mbarrier
//   epilogue  of the PREVIOUS tile while the MMAs of this one run: thread = (row, half of the states): tcgen05.ld,
//             + ck', row maximum (exchanged between the two halves through shared memory), second pass
//             b = exp(ll - max) (float32) written row-major, mx (float64)
#pragma once
#include <cuda.h>
#include "stats_tc.cuh"

#define EDT_NT 256
#define EDT_RT 128

struct EdtArgs {
  int B, T, K, D, N;           // N: MMA N = K rounded up to 16 (<= 256)
  int ntpw, ntiles, mask_ll;
  const int64_t* starts; const uint8_t* mask;
  const double* par2; const double* ckp;
  float* bout; double* mx;
};

struct EdtSmem { size_t Bh, Bl, Ah, Al, xs, ck, pmax, dead, bars, total; };
__host__ __device__ inline EdtSmem edt_layout(int D, int N) {
  EdtSmem s;
  const size_t kd = 2 * (size_t)D;                    // contraction length
  const size_t bsz = (size_t)N * kd * 2, asz = (size_t)EDT_RT * kd * 2;
  s.Bh = 0; s.Bl = bsz; s.Ah = 2 * bsz; s.Al = s.Ah + asz;
  s.xs = s.Al + asz;
  s.ck = s.xs + (size_t)EDT_RT * D * 4;
  s.pmax = s.ck + 256 * 4;
  s.dead = s.pmax + 2 * EDT_RT * 4;
  s.bars = s.dead + 2 * EDT_RT;
  s.total = s.bars + 4 * 8;                          // 226.3 KB at K = 256, D = 64: no static shared memory, no slack -
  return s;                                          // the kernel traps if its dynamic window is not 1024-byte aligned
}

__device__ __forceinline__ void edt_ld32(const uint32_t ta, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                 "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                 "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
                 "=r"(v[30]), "=r"(v[31]) : "r"(ta) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(EDT_NT, 1)
k_emit_diag_tc(const __grid_constant__ CUtensorMap tm_x, const EdtArgs a) {
  extern __shared__ __align__(1024) uint8_t edt_raw[];
  uint8_t* sm = edt_raw;
  if ((dn_smem(edt_raw) & 1023u) != 0) __trap();      // SWIZZLE_128B atoms need 1024-byte alignment (see edt_layout)
  const int K = a.K, D = a.D, T = a.T, N = a.N, KD = 2 * D;
  const EdtSmem L = edt_layout(D, N);
  uint8_t* sBh = sm + L.Bh; uint8_t* sBl = sm + L.Bl; uint8_t* sAh = sm + L.Ah; uint8_t* sAl = sm + L.Al;
  float* xs = reinterpret_cast<float*>(sm + L.xs);
  float* sck = reinterpret_cast<float*>(sm + L.ck);
  float* pmax = reinterpret_cast<float*>(sm + L.pmax);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + L.bars);
  unsigned long long* x_full = bars;                  // observation tile landed
  unsigned long long* mma_done = bars + 1;            // [2]: the MMAs into TMEM stage 0 / 1
  uint8_t* deadrow = sm + L.dead;                     // [2][128] by tile parity: rows without evidence (NaN / masked / past the window)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < 3; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dn_smem(bars + i)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x) : "memory");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dn_smem(tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int k = tid; k < 256; k += EDT_NT) sck[k] = k < K ? (float)a.ckp[k] : 0.f;
  // B operand: state n, contraction element e: e < D -> c2[e][n], else c1[e - D][n]  (par2 is [d][K] pairs)
  for (int task = tid; task < N * (KD / 8); task += EDT_NT) {
    const int n = task % N, c = task / N;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = 8 * c + j, d = e < D ? e : e - D;
      v[j] = n < K ? (float)a.par2[2 * ((size_t)d * K + n) + (e < D ? 0 : 1)] : 0.f;
    }
    uint4 hi, lo;
    stc_pack(v, hi, lo);
    const uint32_t off = dn_chunk(n, c, N);
    *reinterpret_cast<uint4*>(sBh + off) = hi; *reinterpret_cast<uint4*>(sBl + off) = lo;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = *tmem_slot;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const unsigned xbytes = (unsigned)EDT_RT * D * 4;
  const int ntl = a.ntiles > (int)blockIdx.x ? (a.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  auto tile_w = [&](const int j) { return (int)((blockIdx.x + (unsigned)j * gridDim.x) / (unsigned)a.ntpw); };
  auto tile_t0 = [&](const int j) { return (int)((blockIdx.x + (unsigned)j * gridDim.x) % (unsigned)a.ntpw) * EDT_RT; };
  if (tid == 0 && ntl > 0) stc_tma_2d(&tm_x, xs, x_full, 0, (int)(a.starts[tile_w(0)] + tile_t0(0)), xbytes);
  const int row = tid & 127, half = tid >> 7;
  const uint32_t tlane = (uint32_t)((wp & 3) * 32) << 16;
  const int ncol = N / 2;                             // columns of this thread's half (multiple of 8)
  // epilogue of tile j (TMEM stage j & 1)
  auto epilogue = [&](const int j) {
    const int w = tile_w(j), t0 = tile_t0(j);
    const int nrow = min(EDT_RT, T - t0);
    stc_wait(mma_done + (j & 1), (j >> 1) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const bool dead = deadrow[(j & 1) * EDT_RT + row] != 0;
    const uint32_t ta = tm + tlane + (uint32_t)(j & 1) * 256u + (uint32_t)half * ncol;
    float m = -INFINITY;
    for (int c0 = 0; c0 < ncol; c0 += 32) {
      uint32_t v[32];
      edt_ld32(ta + c0, v);
#pragma unroll
      for (int u = 0; u < 32; ++u) {
        const int k = half * ncol + c0 + u;
        if (c0 + u < ncol && k < K) m = fmaxf(m, __uint_as_float(v[u]) + sck[k]);
      }
    }
    if (dead) m = 0.f;
    pmax[half * EDT_RT + row] = m;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    m = fmaxf(pmax[row], pmax[EDT_RT + row]);
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const size_t grow = (size_t)w * T + t0 + row;
    float* bp = a.bout + grow * K;
    for (int c0 = 0; c0 < ncol; c0 += 32) {
      uint32_t v[32];
      edt_ld32(ta + c0, v);
      if (row < nrow) {
#pragma unroll
        for (int u = 0; u < 32; u += 4) {
          const int k = half * ncol + c0 + u;
          if (c0 + u < ncol && k + 3 < K) {
            float4 o;
            o.x = dead ? 1.f : __expf(__uint_as_float(v[u]) + sck[k] - m);
            o.y = dead ? 1.f : __expf(__uint_as_float(v[u + 1]) + sck[k + 1] - m);
            o.z = dead ? 1.f : __expf(__uint_as_float(v[u + 2]) + sck[k + 2] - m);
            o.w = dead ? 1.f : __expf(__uint_as_float(v[u + 3]) + sck[k + 3] - m);
            *reinterpret_cast<float4*>(bp + k) = o;
          }
        }
      }
    }
    if (half == 0 && row < nrow) a.mx[grow] = (double)m;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  };
  for (int j = 0; j < ntl; ++j) {
    const int w = tile_w(j), t0 = tile_t0(j);
    // the MMAs of tile j - 1 read the A operand: done before it is rewritten (their epilogue ran in iteration j - 1
    // only for tile j - 2, so wait here without consuming the phase: the epilogue below waits on it again)
    if (j > 0) { stc_wait(mma_done + ((j - 1) & 1), ((j - 1) >> 1) & 1); asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
    stc_wait(x_full, j & 1);
    // ---- A operand: thread = (row, half of the D dimensions): features x^2 at [0, D), x at [D, 2D)
    {
      const int d0 = half * (D / 2);
      bool bad = false;
      for (int i = 0; i < D / 4; ++i) {                // the whole row decides whether it carries evidence
        const int d = ((i + row) % (D / 4)) * 4;       // (rotated by row: thread = row reads would hit one bank)
        const float4 q = *reinterpret_cast<const float4*>(xs + row * D + d);
        bad |= !(fabsf(q.x) <= 3.0e38f) | !(fabsf(q.y) <= 3.0e38f) | !(fabsf(q.z) <= 3.0e38f) | !(fabsf(q.w) <= 3.0e38f);
      }
      if (t0 + row >= T) bad = true;
      if (!bad && a.mask_ll && a.mask && a.mask[a.starts[w] + t0 + row]) bad = true;
      if (half == 0) deadrow[(j & 1) * EDT_RT + row] = bad ? 1 : 0;
      for (int c8i = 0; c8i < D / 16; ++c8i) {
        const int c8 = (c8i + (row >> 1)) % (D / 16);
        float x[8], x2[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { x[u] = bad ? 0.f : xs[row * D + d0 + 8 * c8 + u]; x2[u] = x[u] * x[u]; }
        uint4 hi, lo;
        const int c = (d0 >> 3) + c8;                  // 16-byte chunk of the x^2 block; + D / 8: the x block
        stc_pack(x2, hi, lo);
        *reinterpret_cast<uint4*>(sAh + dn_chunk(row, c, EDT_RT)) = hi; *reinterpret_cast<uint4*>(sAl + dn_chunk(row, c, EDT_RT)) = lo;
        stc_pack(x, hi, lo);
        *reinterpret_cast<uint4*>(sAh + dn_chunk(row, c + D / 8, EDT_RT)) = hi; *reinterpret_cast<uint4*>(sAl + dn_chunk(row, c + D / 8, EDT_RT)) = lo;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
      if (j + 1 < ntl) stc_tma_2d(&tm_x, xs, x_full, 0, (int)(a.starts[tile_w(j + 1)] + tile_t0(j + 1)), xbytes);
      const uint32_t aAh = dn_smem(sAh), aAl = dn_smem(sAl), aBh = dn_smem(sBh), aBl = dn_smem(sBl);
      const uint32_t dcol = tm + (uint32_t)(j & 1) * 256u;
#pragma unroll 1
      for (int ks = 0; ks < KD / 16; ++ks) {
        const uint32_t oa = (ks >> 2) * (EDT_RT * 128) + (ks & 3) * 32, ob = (ks >> 2) * (N * 128) + (ks & 3) * 32;
        const uint64_t dah = dn_desc(aAh + oa), dal = dn_desc(aAl + oa), dbh = dn_desc(aBh + ob), dbl = dn_desc(aBl + ob);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(dcol), "l"(dal), "l"(dbh), "r"(idesc), "r"(ks > 0 ? 1u : 0u) : "memory");
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(dcol), "l"(dah), "l"(dbl), "r"(idesc), "r"(1u) : "memory");
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(dcol), "l"(dah), "l"(dbh), "r"(idesc), "r"(1u) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(dn_smem(mma_done + (j & 1))) : "memory");
    }
    if (j > 0) epilogue(j - 1);                        // under the MMAs of tile j
  }
  if (ntl > 0) epilogue(ntl - 1);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}
