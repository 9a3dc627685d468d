// Full-covariance expected log-likelihoods on the 5th-generation tensor cores (BASELINE config 3:
// K = 64, D = 32; config 5: 128 mixture components, D = 16): replaces k_emit_full_rb (float64 DFMA,
// 1.87 ms of the 3.0 ms step at c3) for float32 series.
//
//   ll[r][k] = ck[k] - | Rs_k x_r - gk_k |^2            (pybasicbayes/distributions.py:351-366, looped
//                                                        over states at hmmsgd_metaobs.py:508-509)
// is the GEMM  Y[r][(k, i)] = sum_d x[r][d] Rs_k[i][d]  followed by a square-and-sum over i.  A float32 or
// 3xTF32 product misses the 1e-5 parity bound by an order of magnitude (the quadratic
// form is the precision bottleneck of the whole path), so the product is made EXACT:
//   slices    x[r][d] = sx_r * sum_s dx_s 2^(-8-9s),  Rs[c][d] = sR_c * sum_s dR_s 2^(-8-9s),  s = 0..3,
//             with power-of-two scales per observation row / per factor row and integer digits
//             |d| <= 256 (round to nearest), stored as float16 (exact).
//   levels    L_l = sum_{i+j=l} X_i R_j^T, l = 0..3: every product of digits is an integer <= 2^16, a
//             level sums at most 4 * 32 of them (< 2^24), so the float32 accumulation in TENSOR MEMORY is
//             exact; 10 slice products (20 tcgen05.mma of M = 128, N = 64, K = 16 at D = 32) per chunk of
//             64 factor rows, the four levels in four accumulators.
//   epilogue  y = sx_r sR_c 2^-16 (L0 + 2^-9 L1 + 2^-18 L2 + 2^-27 L3) - gk, q += y^2 in float64.  The
//             levels are recombined as INTEGERS (float -> float64 conversions run at 16 lanes per clock
//             on this part, scripts/microbench/ubench6.cu): L0 and L1 are read out of the float mantissa
//             by a magic-number add, W = (512 L0 + L1) 256 + round((L2 + 2^-9 L3) / 2) is a 64-bit
//             integer, and W 2^e (e: the row's scale exponent) is built by adding W to the mantissa of
//             1.5 * 2^(52+e): one DADD + two DFMA on the FP64 pipe per factor row instead of D/2 + 1.
// Slices carry 36 bits below the row maximum: the quadratic form agrees with the float64 kernel to
// ~4e-10 relative (scripts/probes/slice_emulation.py).
//
// One persistent CTA per SM walks tiles of 128 consecutive rows of one window; 10 warps:
//   warps 0-7  (two warpgroups): slice the next tile's observations into the A operand (K-major
//              SWIZZLE_128B, [X0 | X1 | X2 | X3] along K), then the epilogue: thread = row (TMEM lane),
//              warpgroup g takes columns [32 g, 32 g + 32) of every chunk (tcgen05.ld), i.e. the states
//              of its half of the state range; the log-likelihoods of the row wait in a thread-private
//              shared-memory column; at the end of the tile the row maximum is exchanged between the two
//              warpgroups and each thread writes its 128 contiguous bytes of b = exp(ll - max)
//   warp 8     TMA producer: cp.async.bulk.tensor.2d of the observation tiles (tensor map over the
//              (T_full, D) series, SWIZZLE_128B / 64B so that thread = row reads are conflict-free) and
//              cp.async.bulk of the factor chunks (operand image + constants, written pre-swizzled by
//              k_etc_prep) into a ring of ETC_RING slots; mbarrier full / empty pairs
//   warp 9     one thread issues the MMAs of a chunk into TMEM stage (chunk & 1) and commits to the
//              stage's mbarrier; the two stages (2 x 4 levels x 64 columns = 512 columns) overlap the
//              MMAs of chunk c + 1 with the epilogue of chunk c.
// Chunk c of the factor blob holds, in its first 32 rows, the factor rows of states [c SPW, +SPW) and in
// its last 32 rows those of states [H + c SPW, +SPW) (SPW = 32 / D, H = half of the state range).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include "stats_tc.cuh"

#define ETC_NT 320
#define ETC_NCOL 64           // factor rows (accumulator columns) per chunk
#define ETC_RING 4
#define ETC_RT 128            // rows per tile

// bytes of one factor chunk in the blob / in a ring slot: operand image (64 rows x 4 D float16) + (cs, gk) pairs
__host__ __device__ constexpr int etc_img(int D) { return ETC_NCOL * 4 * D * 2; }
__host__ __device__ constexpr int etc_slot(int D) { return etc_img(D) + ETC_NCOL * 16; }

struct EtcArgs {
  int B, T, K;                 // K: emission components (states, or states x mixture components)
  int ntpw, ntiles, nchunks;   // tiles per window, tiles, chunks per tile
  int H;                       // states per warpgroup: nchunks * (32 / D)
  int mask_ll;
  int zero;                    // 0 (a value ptxas cannot fold: orders the two phases of the epilogue)
  const int64_t* starts; const uint8_t* mask;
  const uint8_t* blob;         // [nchunks][etc_slot(D)]
  const double* ck;
  float* bout; double* mx;     // b = exp(ll - max) and the row maxima (K <= 64), or
  double* ll;                  // the float64 log-likelihoods themselves (mixtures: any K), bout == nullptr
  long long* dbg;              // SVIHMM_ETC_DBG: clock64 stamps of CTA 0, second tile: [chunk][8]
};

// factor rows -> float16 digit slices in the tcgen05 K-major SWIZZLE_128B image, + per-row constants.
// grid = nchunks, block = 64 (thread = factor row n of the chunk: half h = n / 32, state
// k = h H + chunk SPW + (n % 32) / D, row i = n % D of Rs_k)
template <int D>
__global__ void __launch_bounds__(ETC_NCOL)
k_etc_prep(int K, int H, const double* __restrict__ Rs, const double* __restrict__ gk, uint8_t* __restrict__ blob) {
  constexpr int tri = D * (D + 1) / 2;
  const int n = threadIdx.x;
  const int k = (n >> 5) * H + (int)blockIdx.x * (32 / D) + (n & 31) / D, i = n % D;
  uint8_t* img = blob + (size_t)blockIdx.x * etc_slot(D);
  double* cst = reinterpret_cast<double*>(img + etc_img(D)) + 2 * n;
  double r[D];
  double amax = 0.0;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    r[j] = (k < K && j <= i) ? Rs[(size_t)k * tri + (size_t)i * (i + 1) / 2 + j] : 0.0;
    amax = fmax(amax, fabs(r[j]));
  }
  int e = 0;
  if (amax > 0.0 && amax < 1e300) { (void)frexp(amax, &e); } // amax = m 2^e, m in [0.5, 1): |r| / 2^e < 1
  const double sR = ldexp(1.0, e), inv = ldexp(256.0, -e);
  __half dg[4][D];
#pragma unroll
  for (int j = 0; j < D; ++j) {
    double v = r[j] * inv;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const double d = rint(v);
      dg[s][j] = __float2half_rn((float)d);
      v = (v - d) * 512.0;
    }
  }
#pragma unroll
  for (int s = 0; s < 4; ++s)
#pragma unroll
    for (int c8 = 0; c8 < D / 8; ++c8) {
      const int c = s * (D / 8) + c8;                      // 16-byte chunk (8 elements) along K
      uint4 pk;
      __half2 h0 = __halves2half2(dg[s][8 * c8], dg[s][8 * c8 + 1]), h1 = __halves2half2(dg[s][8 * c8 + 2], dg[s][8 * c8 + 3]);
      __half2 h2 = __halves2half2(dg[s][8 * c8 + 4], dg[s][8 * c8 + 5]), h3 = __halves2half2(dg[s][8 * c8 + 6], dg[s][8 * c8 + 7]);
      pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
      pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
      if (D == 32)   // 128 operand rows: row (s & 1) 64 + n holds R_s in the K-half s >> 1 (see etc_mma_chunk32)
        *reinterpret_cast<uint4*>(img + dn_chunk((s & 1) * ETC_NCOL + n, (s >> 1) * 4 + c8, 2 * ETC_NCOL)) = pk;
      else
        *reinterpret_cast<uint4*>(img + dn_chunk(n, c, ETC_NCOL)) = pk;
    }
  cst[0] = (k < K) ? sR * (1.0 / 8589934592.0) : 0.0;       // sR 2^-33 (see etc_combine)
  cst[1] = (k < K) ? gk[(size_t)k * D + i] : 0.0;
}

struct EtcSmem { size_t A, ring, xs, llt, pmax, ck, bars, total; };
__host__ __device__ inline EtcSmem etc_layout(int D, int K) {
  EtcSmem s;
  const size_t asz = (size_t)ETC_RT * 4 * D * 2;
  s.A = 0;
  s.ring = 2 * asz;
  s.xs = s.ring + (size_t)ETC_RING * etc_slot(D);
  s.llt = s.xs + (size_t)ETC_RT * D * 4;
  s.pmax = s.llt + (size_t)64 * ETC_RT * 8;               // [64 states][128 rows] float64, thread-private slots
  s.ck = s.pmax + 2 * ETC_RT * 8;
  s.bars = s.ck + (size_t)((K + 63) / 64 * 64 + 8) * 8;
  s.total = s.bars + (1 + 2 + 2 * ETC_RING + 2 + 2) * 8 + 1024;   // + slack for the 1024-byte alignment of the base
  return s;
}

__device__ __forceinline__ void etc_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(dn_smem(bar)) : "memory");
}
// producer-side wait: the polling thread shares a scheduler with two epilogue warps, so it sleeps between polls
__device__ __forceinline__ void etc_wait_sleep(unsigned long long* bar, const unsigned parity) {
  unsigned ok = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(dn_smem(bar)), "r"(parity) : "memory");
    if (ok) break;
    __nanosleep(40);
  }
}
// two barriers polled together (their latencies overlap)
__device__ __forceinline__ void etc_wait2(unsigned long long* b0, const unsigned p0, unsigned long long* b1, const unsigned p1) {
  unsigned ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p, q;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 q, [%3], %4;\n\tand.pred p, p, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(dn_smem(b0)), "r"(p0), "r"(dn_smem(b1)), "r"(p1) : "memory");
}
__device__ __forceinline__ void etc_ld8(const uint32_t ta, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(ta) : "memory");
}

// The four exact level sums of one accumulator column (float32 bit patterns) -> W 2^e in float64 with
// W = (512 L0 + L1) 256 + rint((L2 + 2^-9 L3) / 2):  sum_d x_d Rs_cd = sx sR 2^-33 W  (sx = 2^e).
//   L0, L1 (|.| <= 2^22, exact integers): x + 1.5 * 2^23 has the integer in its mantissa;
//   L2 + 2^-9 L3 (|.| < 2^23): + 1.5 * 2^24 leaves half of it, rounded - 2^-39 of the leading level;
//   the 64-bit integer W + 0x4BC00000 (the float bias of the third term is left in) is added to the bit
//   pattern of 1.5 * 2^(52+e) (hi word himag), which is the float64 number 1.5 * 2^(52+e) + (W + bias) 2^e
//   exactly; negmag = -(1.5 * 2^(52+e) + bias 2^e), whose bit pattern is (himag, bias).
// The FP64 pipe of this part does NOT run while tcgen05.mma is executing (scripts/microbench/ubench7.cu:
// a DFMA loop and an MMA stream take the SUM of their times; FFMA is unaffected), and the MMAs of chunk
// c + 1 start exactly when the epilogue of chunk c does.  So the epilogue of a chunk is two phases: all the
// float32 / integer work first (under the MMAs), then all the float64 instructions; a data dependency
// through `negmag` keeps ptxas from interleaving them.
__device__ __forceinline__ void etc_combine(const uint32_t a0, const uint32_t a1, const uint32_t a2, const uint32_t a3,
                                            const int himag, uint32_t& whi, uint32_t& wlo) {
  const uint32_t b0 = __float_as_uint(__uint_as_float(a0) + 12582912.f);
  const uint32_t b1 = __float_as_uint(__uint_as_float(a1) + 12582912.f);
  const float tl = fmaf(__uint_as_float(a3), 0.001953125f, __uint_as_float(a2));
  const uint32_t b2 = __float_as_uint(tl + 25165824.f);                  // 0x4BC00000 + rint(tl / 2)
  const int M = (int)(b0 * 512u + b1 - 0x4B400000u * 513u);              // 512 L0 + L1 (the biases wrap away)
  // one IMAD.WIDE: the addend carries the hi-word magic and the (still biased) low integer
  const long long W = (long long)M * 256 + (long long)(((unsigned long long)(uint32_t)himag << 32) | b2);
  whi = (uint32_t)(W >> 32); wlo = (uint32_t)W;
}

// one tcgen05.mma (float16 in, float32 accumulate); MODE: the A operand's collector use - 0 none, 1 fill (keep
// for the following MMAs), 2 use (same A again), 3 last use.  A slice of X meets up to four slices of Rs in
// consecutive MMAs, so it is fetched from shared memory once (SASS: UTCHMMA ... .A_KEEP / .A_REUSE).
template <int MODE>
__device__ __forceinline__ void etc_mma(const uint32_t dcol, const uint64_t da, const uint64_t db, const uint32_t idesc, const uint32_t acc) {
  if (MODE == 0)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(dcol), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
  else if (MODE == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(dcol), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
  else if (MODE == 2)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16.collector::a::use [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(dcol), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(dcol), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
// all MMAs of one chunk: for every k-step of every slice X_i, the products with R_0 .. R_(3-i) into the level
// accumulators i .. 3 (element offset e along K -> descriptor address offset, in 16-byte units)
template <int D, int I, int KSI>
__device__ __forceinline__ void etc_mma_slice(const uint32_t d0, const uint64_t dA, const uint64_t dB, const uint32_t idesc) {
  constexpr int ea = I * D + KSI * 16;
  constexpr uint32_t offA = ((ea >> 6) * (ETC_RT * 128) + (ea & 63) * 2) >> 4;
  constexpr uint32_t acc = (I > 0 || KSI > 0) ? 1u : 0u;
  constexpr int n = 4 - I;
#define ETC_OFFB(J) (uint32_t)(((((J) * D + KSI * 16) >> 6) * (ETC_NCOL * 128) + (((J) * D + KSI * 16) & 63) * 2) >> 4)
  if (n == 1) etc_mma<0>(d0 + I * ETC_NCOL, dA + offA, dB + ETC_OFFB(0), idesc, acc);
  else {
    etc_mma<1>(d0 + I * ETC_NCOL, dA + offA, dB + ETC_OFFB(0), idesc, acc);
    if (n > 2) etc_mma<2>(d0 + (I + 1) * ETC_NCOL, dA + offA, dB + ETC_OFFB(1), idesc, acc);
    if (n > 3) etc_mma<2>(d0 + (I + 2) * ETC_NCOL, dA + offA, dB + ETC_OFFB(2), idesc, acc);
    etc_mma<3>(d0 + (I + n - 1) * ETC_NCOL, dA + offA, dB + ETC_OFFB(n - 1), idesc, acc);
  }
#undef ETC_OFFB
}

// D = 32: the factor operand of a chunk is stacked along N - operand row (j & 1) 64 + n, K-half j >> 1 holds
// row n of R_j - so that one MMA with N = 128 meets a slice of X with TWO slices of Rs and lands in two
// adjacent level accumulators: 12 MMAs per chunk (8 with N = 128) instead of 20 with N = 64, which ran at
// 72 cycles each (44 % of the tensor pipe's rate: the A operand is re-read per 64 columns).
__device__ __forceinline__ void etc_mma_chunk32(const uint32_t d0, const uint64_t dA, const uint64_t dB) {
  constexpr uint32_t id128 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  constexpr uint32_t id64 = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    const uint32_t acc0 = ks ? 1u : 0u;
    // A: slice i at elements [32 i, 32 i + 32) of the 128-element row (two SWIZZLE_128B atoms of 64)
    const uint64_t a0 = dA + ((0 * 64 + ks * 32) >> 4), a1 = dA + ((1 * 64 + ks * 32) >> 4);
    const uint64_t a2 = dA + ((ETC_RT * 128 + 0 * 64 + ks * 32) >> 4), a3 = dA + ((ETC_RT * 128 + 1 * 64 + ks * 32) >> 4);
    // B: K-half h at bytes [64 h, 64 h + 64) of the 128-byte row
    const uint64_t b0 = dB + ((0 * 64 + ks * 32) >> 4), b1 = dB + ((1 * 64 + ks * 32) >> 4);
    etc_mma<1>(d0 + 0 * ETC_NCOL, a0, b0, id128, acc0);      // X0 [R0; R1] -> levels 0, 1
    etc_mma<3>(d0 + 2 * ETC_NCOL, a0, b1, id128, acc0);      // X0 [R2; R3] -> levels 2, 3
    etc_mma<1>(d0 + 1 * ETC_NCOL, a1, b0, id128, 1u);        // X1 [R0; R1] -> levels 1, 2
    etc_mma<3>(d0 + 3 * ETC_NCOL, a1, b1, id64, 1u);         // X1 R2       -> level 3
    etc_mma<0>(d0 + 2 * ETC_NCOL, a2, b0, id128, 1u);        // X2 [R0; R1] -> levels 2, 3
    etc_mma<0>(d0 + 3 * ETC_NCOL, a3, b0, id64, 1u);         // X3 R0       -> level 3
  }
}

template <int D>
__global__ void __launch_bounds__(ETC_NT, 1)
k_emit_tc(const __grid_constant__ CUtensorMap tm_x, const EtcArgs a) {
  constexpr int SPW = 32 / D;                  // states per chunk and warpgroup
  constexpr int KS = D / 16;                   // MMA k-steps per slice product
  constexpr int ASZ = ETC_RT * 4 * D * 2;      // one A operand
  constexpr int SLOT = etc_slot(D), IMG = etc_img(D);
  extern __shared__ __align__(1024) uint8_t etc_raw[];
  uint8_t* sm = etc_raw + ((1024u - (dn_smem(etc_raw) & 1023u)) & 1023u);
  const EtcSmem L = etc_layout(D, a.K);
  uint8_t* sA = sm + L.A; uint8_t* ring = sm + L.ring;
  uint8_t* xs = sm + L.xs;
  double* llt = reinterpret_cast<double*>(sm + L.llt);
  double* pmax = reinterpret_cast<double*>(sm + L.pmax);
  double* sck = reinterpret_cast<double*>(sm + L.ck);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(sm + L.bars);
  unsigned long long* x_full = bars;                         // observation tile landed (TMA)
  unsigned long long* a_full = bars + 1;                     // [2] A operand written (256 arrivals)
  unsigned long long* b_full = bars + 3;                     // [RING] factor chunk landed
  unsigned long long* b_empty = bars + 3 + ETC_RING;         // [RING] chunk consumed by the epilogue threads (256 arrivals: every reader of the constants releases its own reads)
  unsigned long long* t_full = bars + 3 + 2 * ETC_RING;      // [2] MMAs of the stage complete (tcgen05.commit)
  unsigned long long* t_empty = t_full + 2;                  // [2] stage read back (8 arrivals)
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  const int K = a.K, T = a.T, nch = a.nchunks;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dn_smem(x_full)) : "memory");
    for (int i = 0; i < 2; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 256;" ::"r"(dn_smem(a_full + i)) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dn_smem(t_full + i)) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(dn_smem(t_empty + i)) : "memory");
    }
    for (int i = 0; i < ETC_RING; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dn_smem(b_full + i)) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 256;" ::"r"(dn_smem(b_empty + i)) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_x) : "memory");
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dn_smem(&tmem_base)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int k = tid; k < (K + 63) / 64 * 64 + 8; k += ETC_NT) sck[k] = k < K ? a.ck[k] : 0.0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_base;
  const int ntl = a.ntiles > (int)blockIdx.x ? (a.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;   // tiles of this CTA
  auto tile_w = [&](const int j) { return (int)((blockIdx.x + (unsigned)j * gridDim.x) / (unsigned)a.ntpw); };
  auto tile_t0 = [&](const int j) { return (int)((blockIdx.x + (unsigned)j * gridDim.x) % (unsigned)a.ntpw) * ETC_RT; };
  constexpr unsigned xbytes = ETC_RT * D * 4;

  if (wp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0 && ntl > 0) {
      stc_tma_2d(&tm_x, xs, x_full, 0, (int)(a.starts[tile_w(0)] + tile_t0(0)), xbytes);
      unsigned g = 0;
      for (int j = 0; j < ntl; ++j) {
        // tile j has been sliced (one iteration ahead of its MMAs): the staging buffer takes tile j + 1
        etc_wait_sleep(a_full + (j & 1), (j >> 1) & 1);
        if (j + 1 < ntl)
          stc_tma_2d(&tm_x, xs, x_full, 0, (int)(a.starts[tile_w(j + 1)] + tile_t0(j + 1)), xbytes);
        for (int c = 0; c < nch; ++c, ++g) {
          const unsigned slot = g % ETC_RING, use = g / ETC_RING;
          if (use > 0) etc_wait_sleep(b_empty + slot, (use - 1) & 1);
          if (a.dbg && blockIdx.x == 0 && j == 1) a.dbg[c * 8 + 0] = clock64();
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dn_smem(b_full + slot)), "r"(SLOT) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(dn_smem(ring + (size_t)slot * SLOT)), "l"(a.blob + (size_t)c * SLOT), "r"(SLOT), "r"(dn_smem(b_full + slot)) : "memory");
        }
      }
    }
  } else if (wp == 9) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0 && ntl > 0) {
      // float16 inputs (formats 0), float32 accumulators, M = 128, N = 64
      const uint32_t idesc = (1u << 4) | ((uint32_t)(ETC_NCOL >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      unsigned g = 0;
      for (int j = 0; j < ntl; ++j) {
        etc_wait_sleep(a_full + (j & 1), (j >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t dA = dn_desc(dn_smem(sA + (size_t)(j & 1) * ASZ));
        for (int c = 0; c < nch; ++c, ++g) {
          const unsigned slot = g % ETC_RING, stage = g & 1, use = g >> 1;
          etc_wait_sleep(b_full + slot, (g / ETC_RING) & 1);
          if (a.dbg && blockIdx.x == 0 && j == 1) a.dbg[c * 8 + 1] = clock64();
          if (use > 0) etc_wait_sleep(t_empty + stage, (use - 1) & 1);
          if (a.dbg && blockIdx.x == 0 && j == 1) a.dbg[c * 8 + 2] = clock64();
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t dB = dn_desc(dn_smem(ring + (size_t)slot * SLOT));
          const uint32_t d0 = tm + stage * 256u;
          if (D == 32) etc_mma_chunk32(d0, dA, dB);
          else {
          etc_mma_slice<D, 0, 0>(d0, dA, dB, idesc); if (KS > 1) etc_mma_slice<D, 0, KS - 1>(d0, dA, dB, idesc);
          etc_mma_slice<D, 1, 0>(d0, dA, dB, idesc); if (KS > 1) etc_mma_slice<D, 1, KS - 1>(d0, dA, dB, idesc);
          etc_mma_slice<D, 2, 0>(d0, dA, dB, idesc); if (KS > 1) etc_mma_slice<D, 2, KS - 1>(d0, dA, dB, idesc);
          etc_mma_slice<D, 3, 0>(d0, dA, dB, idesc); if (KS > 1) etc_mma_slice<D, 3, KS - 1>(d0, dA, dB, idesc);
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(dn_smem(t_full + stage)) : "memory");
          if (a.dbg && blockIdx.x == 0 && j == 1) a.dbg[c * 8 + 3] = clock64();
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ slicers + epilogue (256 threads)
    const int row = tid & 127, wg = tid >> 7;
    // observation tile j -> A operand (j & 1); returns this row's scale (0: the row carries no evidence)
    auto slice = [&](const int j) -> float {
      const int w = tile_w(j), t0 = tile_t0(j);
      stc_wait(x_full, j & 1);
      const uint8_t* xb = xs;
      float x[D];
#pragma unroll
      for (int q = 0; q < D / 4; ++q) {
        const int pq = D == 32 ? (q ^ (row & 7)) : (q ^ ((row >> 1) & 3));
        const float4 v = *reinterpret_cast<const float4*>(xb + (size_t)row * D * 4 + pq * 16);
        x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
      }
      float amax = 0.f; bool dead = t0 + row >= T;
#pragma unroll
      for (int d = 0; d < D; ++d) { amax = fmaxf(amax, fabsf(x[d])); dead |= !(fabsf(x[d]) <= 1.0e30f); }   // NaN / inf / beyond the scale range
      if (!dead && a.mask_ll && a.mask && a.mask[a.starts[w] + t0 + row]) dead = true;
      unsigned e = (__float_as_uint(amax) >> 23) & 0xffu;
      e = e < 8u ? 8u : e;                                       // zeros / tiny rows: scale 2^-118
      const float sx = __uint_as_float((e + 1u) << 23);          // 2^(e+1-127) > amax
      const float inv = __uint_as_float((254u - (e + 1u) + 8u) << 23);   // 256 / sx
      uint8_t* A = sA + (size_t)(j & 1) * ASZ;
      constexpr int NV = D / 2;                                  // values of this thread: [wg * NV, +NV)
      __half dg[4][NV];
#pragma unroll
      for (int u = 0; u < NV; ++u) {
        float v = dead ? 0.f : (wg ? x[NV + u] : x[u]) * inv;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const float d = rintf(v);
          dg[s][u] = __float2half_rn(d);
          v = (v - d) * 512.f;
        }
      }
#pragma unroll
      for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int c8 = 0; c8 < NV / 8; ++c8) {
          const int c = s * (D / 8) + wg * (NV / 8) + c8;
          uint4 pk;
          __half2 h0 = __halves2half2(dg[s][8 * c8], dg[s][8 * c8 + 1]), h1 = __halves2half2(dg[s][8 * c8 + 2], dg[s][8 * c8 + 3]);
          __half2 h2 = __halves2half2(dg[s][8 * c8 + 4], dg[s][8 * c8 + 5]), h3 = __halves2half2(dg[s][8 * c8 + 6], dg[s][8 * c8 + 7]);
          pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
          pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
          *reinterpret_cast<uint4*>(A + dn_chunk(row, c, ETC_RT)) = pk;
        }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      etc_arrive(a_full + (j & 1));
      return dead ? 0.f : sx;
    };
    float sx_cur = 0.f, sx_next = 0.f;
    if (ntl > 0) sx_cur = slice(0);
    const uint32_t tlane = (uint32_t)((wp & 3) * 32) << 16;
    double* myll = llt + row;                                    // [state of this warpgroup][128 rows]: thread-private
    const int H = a.H;
    for (int j = 0; j < ntl; ++j) {
      if (j + 1 < ntl) sx_next = slice(j + 1);
      const int w = tile_w(j), t0 = tile_t0(j);
      const int nrow = min(ETC_RT, T - t0);
      const bool dead = sx_cur == 0.f;
      // per-row constants of etc_combine: sx = 2^e
      const int e = (int)((__float_as_uint(sx_cur) >> 23) & 0xffu) - 127;
      const int himag = 0x43380000 + (dead ? 0 : e) * 0x100000;
      const double negmag = -__hiloint2double(himag, 0x4BC00000);
      const size_t grow = (size_t)w * T + t0 + row;
#pragma unroll 1
      for (int c = 0; c < nch; ++c) {
        const unsigned g = (unsigned)j * nch + c, slot = g % ETC_RING, stage = g & 1;
        const bool stamp = a.dbg && blockIdx.x == 0 && j == 1 && tid == 0;
        if (stamp) a.dbg[c * 8 + 4] = clock64();
        etc_wait2(t_full + stage, (g >> 1) & 1, b_full + slot, (g / ETC_RING) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (stamp) a.dbg[c * 8 + 5] = clock64();
        const double2* cst = reinterpret_cast<const double2*>(ring + (size_t)slot * SLOT + IMG) + wg * 32;
        // phase 1 (overlaps the MMAs of the next chunk): the accumulator columns, 8 at a time (four levels
        // each; the loads of group cg + 1 are in flight while group cg is reduced) -> 64-bit integers
        uint32_t v[2][4][8], whi[32], wlo[32];
        const uint32_t ta = tm + tlane + stage * 256u + (uint32_t)wg * 32u;
        etc_ld8(ta, v[0][0]); etc_ld8(ta + ETC_NCOL, v[0][1]); etc_ld8(ta + 2 * ETC_NCOL, v[0][2]); etc_ld8(ta + 3 * ETC_NCOL, v[0][3]);
#pragma unroll
        for (int cg = 0; cg < 4; ++cg) {
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (cg + 1 < 4) {
            const uint32_t tn = ta + (cg + 1) * 8;
            etc_ld8(tn, v[(cg + 1) & 1][0]); etc_ld8(tn + ETC_NCOL, v[(cg + 1) & 1][1]);
            etc_ld8(tn + 2 * ETC_NCOL, v[(cg + 1) & 1][2]); etc_ld8(tn + 3 * ETC_NCOL, v[(cg + 1) & 1][3]);
          }
#pragma unroll
          for (int u = 0; u < 8; ++u)
            etc_combine(v[cg & 1][0][u], v[cg & 1][1][u], v[cg & 1][2][u], v[cg & 1][3][u], himag, whi[cg * 8 + u], wlo[cg * 8 + u]);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (stamp) a.dbg[c * 8 + 6] = clock64();
        // phase 2: float64.  nm depends on every integer of phase 1 (a.zero is 0 at run time)
        uint32_t dep = 0;
#pragma unroll
        for (int i = 0; i < 32; ++i) dep ^= wlo[i];
        const double nm = __hiloint2double(__double2hiint(negmag), __double2loint(negmag) ^ (int)(dep & (uint32_t)a.zero));
        double acc[SPW][2];
#pragma unroll
        for (int s = 0; s < SPW; ++s) acc[s][0] = acc[s][1] = 0.0;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const double vd = __hiloint2double((int)whi[i], (int)wlo[i]) + nm;
          const double2 cu = cst[i];
          const double y = fma(vd, cu.x, -cu.y);
          acc[i / D][i & 1] = fma(y, y, acc[i / D][i & 1]);
        }
        // this thread's states of the chunk: [k0, k0 + SPW).  Every float64 instruction of the chunk comes
        // BEFORE the stage is handed back (the stores below are ordered before the arrivals)
        const int kl = c * SPW, k0 = wg * H + kl;
        if (a.bout) {
#pragma unroll
          for (int s = 0; s < SPW; ++s) {
            const double vv = k0 + s < K ? (dead ? 0.0 : sck[k0 + s] - (acc[s][0] + acc[s][1])) : -INFINITY;
            myll[(size_t)(wg * 32 + kl + s) * ETC_RT] = vv;
          }
        } else if (row < nrow) {
          double* lp = a.ll + grow * K + k0;
#pragma unroll
          for (int s = 0; s < SPW; ++s)
            if (k0 + s < K) lp[s] = dead ? 0.0 : sck[k0 + s] - (acc[s][0] + acc[s][1]);
        }
        // Only now is the stage handed back (the accumulators have been in registers since phase 1): the MMAs
        // of chunk c + 2 must run under phase 1 of chunk c + 1, not under the float64 phase of this chunk.
        // The constants have been read: the ring slot is free again.
        etc_arrive(b_empty + slot);
        __syncwarp();
        if (lane == 0) etc_arrive(t_empty + stage);
        if (stamp) a.dbg[c * 8 + 7] = clock64();
      }
      if (a.bout) {
        // row maximum over both warpgroups, then this thread's H contiguous values of b = exp(ll - max)
        double pm = -INFINITY;
        for (int k = 0; k < H; ++k) pm = fmax(pm, myll[(size_t)(wg * 32 + k) * ETC_RT]);
        pmax[wg * ETC_RT + row] = pm;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const double m = fmax(pmax[row], pmax[ETC_RT + row]);
        asm volatile("bar.sync 1, 256;" ::: "memory");          // pmax is rewritten at the end of the next tile
        if (row < nrow) {
          float* bp = a.bout + grow * K + wg * H;
          const int nk = min(H, K - wg * H);                     // multiple of 4 (K % 8 == 0, H % 4 == 0)
          for (int k = 0; k < nk; k += 4) {
            const double* lp = myll + (size_t)(wg * 32 + k) * ETC_RT;
            *reinterpret_cast<float4*>(bp + k) = make_float4(__expf((float)(lp[0] - m)), __expf((float)(lp[ETC_RT] - m)),
                                                            __expf((float)(lp[2 * ETC_RT] - m)), __expf((float)(lp[3 * ETC_RT] - m)));
          }
          if (wg == 0) a.mx[grow] = m;
        }
      }
      sx_cur = sx_next;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}
