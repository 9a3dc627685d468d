// k_chain_dense split over a thread-block CLUSTER (K = 256, BASELINE config 4): the one-CTA-per-(tile,
// direction) kernel of dense.cuh launched 16 CTAs on 148 SMs for 1024 windows and spent 10 k cycles per step
// (16 MMAs of N = 256, then a 64-values-per-thread epilogue that cannot start before them).  Here the N (target
// state) dimension of a tile's step is split over the 4 CTAs of a cluster:
//   CTA r holds rows [64 r, 64 r + 64) of the transition operand (32 KB) and the WHOLE carried vector of the 128
//   windows (A operand, 64 KB); per step it issues 16 tcgen05.mma of M = 128, N = 64, K = 16, its 512 threads
//   (4 per window) take 16 accumulator columns each (tcgen05.ld), multiply by b[t] and the exact power-of-two
//   rescaling, store the float32 message, and write the bf16 rounding of their 16 states into the CTA's own slice
//   of the A operand - in the K-major SWIZZLE_128B layout the 64 states of CTA r are exactly atom column r, 16 KB
//   contiguous - with their partial row sums (2 KB).  One thread then pushes slice + sums into the same place of the
//   three peers with cp.async.bulk shared::cta -> shared::cluster, completing on the RECEIVER's mbarrier: the arrival
//   of the three slices is the only synchronisation of a step (no barrier.cluster; a first version with 6144
//   st.shared::cluster per CTA and step + barrier.cluster spent 4-5 k cycles per step there).  The operand and the
//   sums are double buffered by step parity (issuing the MMAs of each source slice as soon as it lands, with one
//   barrier per source CTA, was measured and made no difference: the MMAs are not on the critical path): a peer's slice for step s + 2 can only arrive after this CTA has
//   published step s + 1, i.e. after its MMAs of step s + 1 - the readers of that buffer - are complete.
//   64 CTAs for 1024 windows, a quarter of the epilogue and of the MMA time per step.
// Same tables as k_chain_dense (tile layout; alphaT scaled by 2^-E, E in ET).
#pragma once
#include <cooperative_groups.h>
#include "dense.cuh"

#define DNC_R 4                 // CTAs per cluster = N split
#define DNC_KP 256
#define DNC_NS (DNC_KP / DNC_R) // target states per CTA

__device__ __forceinline__ void dnc_ld16(const uint32_t ta, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                 "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(ta) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// grid (tiles * 4, 2 directions), cluster (4, 1, 1); 512 threads
__global__ void __cluster_dims__(DNC_R, 1, 1) __launch_bounds__(DN_M * 4, 1)
k_chain_dense_cl(int B, int T, int K, const float* __restrict__ Pfwd, const float* __restrict__ Pbwd,
                 const float* __restrict__ pi0, const float* __restrict__ bT, float* __restrict__ alphaT,
                 float* __restrict__ betaT, int* __restrict__ ET) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(1024) uint8_t dnc_raw[];
  uint8_t* dsm = dnc_raw + ((1024u - (dn_smem(dnc_raw) & 1023u)) & 1023u);
  uint8_t* sA = dsm;                                    // [2] DN_M x 256 bf16 (64 KB each): the carried vectors, all states,
                                                        // by step parity (a peer's MMAs may still read the old one)
  uint8_t* sB = dsm + 2 * (size_t)DN_M * DNC_KP * 2;    // DNC_NS x 256 bf16 (32 KB): this CTA's rows of the operand
  float* psum = reinterpret_cast<float*>(sB + (size_t)DNC_NS * DNC_KP * 2);   // [2][16][DN_M] partial row sums by step parity
  __shared__ __align__(8) unsigned long long bar;
  __shared__ __align__(8) unsigned long long recv[2];   // the peers' slices of parity 0 / 1 have landed (3 x 18 KB)
  __shared__ uint32_t tmem_base;
  const int rank = (int)cluster.block_rank();
  const int tid = threadIdx.x & (DN_M - 1), wp = tid >> 5, part = threadIdx.x >> 7;
  const bool fwd = blockIdx.y == 0;
  const int tile = blockIdx.x / DNC_R;
  const float* Mrow = fwd ? Pfwd : Pbwd;
  // B operand: row n = target state 64 rank + nl, column k (zero padded)
  for (int i = threadIdx.x; i < DNC_NS * (DNC_KP / 8); i += DN_M * 4) {
    const int nl = i / (DNC_KP / 8), c = i - nl * (DNC_KP / 8), n = rank * DNC_NS + nl;
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { const int k = 8 * c + u; v[u] = (n < K && k < K) ? __ldg(Mrow + (size_t)n * K + k) : 0.f; }
    *reinterpret_cast<uint4*>(sB + dn_chunk(nl, c, DNC_NS)) = make_uint4(dn_pack(v[0], v[1]), dn_pack(v[2], v[3]), dn_pack(v[4], v[5]), dn_pack(v[6], v[7]));
  }
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dn_smem(&bar)) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dn_smem(&recv[0])) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dn_smem(&recv[1])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dn_smem(&tmem_base)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // this thread's 16 states: k0 + j, j < 16; their two 16-byte chunks of the A operand rows: c0, c0 + 1
  const int k0 = rank * DNC_NS + part * 16, c0 = k0 >> 3, slot = rank * 4 + part;
  constexpr uint32_t ABUF = DN_M * DNC_KP * 2, SLICE = DN_M * 128, PSB = 4 * DN_M * 4;   // operand buffer, slice, 4 sum rows
  uint32_t peerA[DNC_R], peerP[DNC_R], peerBar[DNC_R];  // shared::cluster addresses of sA, psum, recv[0] in every CTA
#pragma unroll
  for (int r = 0; r < DNC_R; ++r) {
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peerA[r]) : "r"(dn_smem(sA)), "r"(r));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peerP[r]) : "r"(dn_smem(psum)), "r"(r));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(peerBar[r]) : "r"(dn_smem(&recv[0])), "r"(r));
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster.sync();                                       // everybody's barriers exist before remote traffic
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // the 16 values of a step: bf16 into this CTA's slice of operand buffer `par`, + the partial row sum; then one
  // thread pushes slice and sums to the three peers
  auto publish = [&](const float (&v)[16], const float sum, const int par) {
    const uint4 q0 = make_uint4(dn_pack(v[0], v[1]), dn_pack(v[2], v[3]), dn_pack(v[4], v[5]), dn_pack(v[6], v[7]));
    const uint4 q1 = make_uint4(dn_pack(v[8], v[9]), dn_pack(v[10], v[11]), dn_pack(v[12], v[13]), dn_pack(v[14], v[15]));
    uint8_t* dstA = sA + (size_t)par * ABUF;
    *reinterpret_cast<uint4*>(dstA + dn_chunk(tid, c0, DN_M)) = q0; *reinterpret_cast<uint4*>(dstA + dn_chunk(tid, c0 + 1, DN_M)) = q1;
    psum[(par * 16 + slot) * DN_M + tid] = sum;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t offA = (uint32_t)par * ABUF + (uint32_t)rank * SLICE, offP = (uint32_t)((par * 16 + rank * 4) * DN_M * 4);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dn_smem(&recv[par])), "r"((DNC_R - 1) * (SLICE + PSB)) : "memory");
#pragma unroll
      for (int r = 0; r < DNC_R; ++r) {
        if (r == rank) continue;
        const uint32_t rb = peerBar[r] + (uint32_t)par * 8;
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(peerA[r] + offA), "r"(dn_smem(sA) + offA), "r"(SLICE), "r"(rb) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(peerP[r] + offP), "r"(dn_smem(psum) + offP), "r"(PSB), "r"(rb) : "memory");
      }
    }
  };
  auto wait_recv = [&](const int par, const unsigned parity) {
    unsigned ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(dn_smem(&recv[par])), "r"(parity) : "memory");
  };
  // step 0: forward alpha_0 = pi0 * b_0; backward beta_{T-1} = 1, carried b_{T-1}
  const int dt = fwd ? 1 : -1;
  int t = fwd ? 0 : T - 1;
  const size_t tb = dn_tile_off(tile, T, K, 0) + tid;
  float* outp = (fwd ? alphaT : betaT) + tb;
  const float* bw = bT + tb;
  int* Ep = ET + (size_t)tile * T * DN_M + tid;
  {
    float v[16], sum = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int k = k0 + j;
      const float bb = (k < K) ? bw[((size_t)t * K + k) * DN_M] : 0.f;
      v[j] = fwd ? ((k < K) ? __ldg(pi0 + k) * bb : 0.f) : bb;
      if (k < K) outp[((size_t)t * K + k) * DN_M] = fwd ? v[j] : 1.f;
      sum += v[j];
    }
    publish(v, sum, 0);
  }
  int E = 0;
  if (fwd && rank == 0 && part == 0) Ep[(size_t)t * DN_M] = 0;
  const uint32_t tm = tmem_base;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(DNC_NS >> 3) << 17) | ((uint32_t)(DN_M >> 4) << 24);
  const uint64_t dA0 = dn_desc(dn_smem(sA)), dA1 = dn_desc(dn_smem(sA + (size_t)DN_M * DNC_KP * 2)), dB = dn_desc(dn_smem(sB));
  uint32_t phase = 0;
  // The messages of a step are stored to global memory at the BEGINNING of the next step, under its MMAs: the
  // release of barrier.cluster waits for every earlier store of the thread to be performed, and with the 16 global
  // stores right in front of it every step paid their L2 round trip (measured: 11 k cycles per step).
  float pend[16]; float* pend_ot = nullptr; int pend_E = 0; int* pend_Ep = nullptr;
  // b of the NEXT step is loaded one whole step ahead (ncu: the kernel sat on long-scoreboard stalls, 7.4 warps per
  // issue, at the first use of b when it was loaded under the MMAs of the same step)
  float bq[16];
  if (T > 1) {
#pragma unroll
    for (int j = 0; j < 16; ++j) bq[j] = (k0 + j < K) ? bw[((size_t)(t + dt) * K + k0 + j) * DN_M] : 0.f;
  }
  for (int s = 1; s < T; ++s) {
    t += dt;
    wait_recv((s - 1) & 1, ((s - 1) >> 1) & 1);        // the carried vectors of step s - 1 are complete in this CTA
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {
      const uint64_t dA = ((s - 1) & 1) ? dA1 : dA0;
#pragma unroll
      for (int ks = 0; ks < DNC_KP / 16; ++ks) {
        const uint64_t da = dA + (((ks >> 2) * (DN_M * 128) + (ks & 3) * 32) >> 4);
        const uint64_t db = dB + (((ks >> 2) * (DNC_NS * 128) + (ks & 3) * 32) >> 4);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(ks > 0 ? 1u : 0u) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(dn_smem(&bar)) : "memory");
    }
    // row sum of the previous step: the 16 partials in a fixed order (the same bits in all 16 threads of a window)
    float sum = 0.f;
    const float* ps = psum + ((s - 1) & 1) * 16 * DN_M + tid;
#pragma unroll
    for (int i = 0; i < 16; ++i) sum += ps[i * DN_M];
    int d = (int)((__float_as_uint(sum) >> 23) & 0xff) - 127;
    if (!(sum > 0.f)) d = 0;
    d = max(-100, min(100, d));
    const float r = __uint_as_float((unsigned)(127 - d) << 23);
    E += d;
    const float* bt = bw + (size_t)t * K * DN_M;
    float* ot = outp + (size_t)t * K * DN_M;
    if (s + 2 < T) {                                   // this CTA's quarter of the b tile of step s + 2 into L2
      const char* pf = reinterpret_cast<const char*>(bT + dn_tile_off(tile, T, K, t + 2 * dt) + (size_t)rank * DNC_NS * DN_M) + (size_t)threadIdx.x * 64;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
    }
    if (pend_ot) {
#pragma unroll
      for (int j = 0; j < 16; ++j) if (k0 + j < K) pend_ot[(size_t)(k0 + j) * DN_M] = pend[j];
      if (pend_Ep) *pend_Ep = pend_E;
    }
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(dn_smem(&bar)), "r"(phase) : "memory");
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t acc[16];
    dnc_ld16(tm + ((uint32_t)(wp * 32) << 16) + part * 16, acc);
    float v[16], s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float m = __uint_as_float(acc[j]) * r;
      v[j] = m * bq[j];
      pend[j] = fwd ? v[j] : m;
      if (j & 1) s1 += v[j]; else s0 += v[j];
    }
    if (s + 1 < T) {
#pragma unroll
      for (int j = 0; j < 16; ++j) bq[j] = (k0 + j < K) ? bw[((size_t)(t + dt) * K + k0 + j) * DN_M] : 0.f;
    }
    pend_ot = ot; pend_E = E; pend_Ep = (fwd && rank == 0 && part == 0) ? Ep + (size_t)t * DN_M : nullptr;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (s + 1 < T) publish(v, s0 + s1, s & 1);
  }
  cluster.sync();                                       // nobody leaves while a peer's bulk copy may still read its slice
  if (pend_ot) {
#pragma unroll
    for (int j = 0; j < 16; ++j) if (k0 + j < K) pend_ot[(size_t)(k0 + j) * DN_M] = pend[j];
    if (pend_Ep) *pend_Ep = pend_E;
  }
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64) : "memory");
}
