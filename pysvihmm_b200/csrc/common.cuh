// Shared declarations for libsvihmm.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include <vector>

#include "../../include/svihmm.h"

#define SVIHMM_EPS 1e-9      /* hmmbase.py:30 / hmmsgd_metaobs.py:26 */
#define SVIHMM_NSLOT 3
#define SVIHMM_WEPS 1e-12    /* pybasicbayes/distributions.py:22     */

struct svihmm_ctx {
  int device, K, D, kind, KP;
  int OD;         // columns of the observation series: D, or 1 for categorical symbols
  int C, KE;      // mixture components per state (1 = plain emissions) and emission components K*C
  size_t plen;    // doubles per state in the packed emission parameters
  size_t slen;    // doubles in the packed statistics
  int nfeat;      // columns of the statistics contraction: K + 1 + D + DD
  int DD;         // D*D (full) or D (diag)
  // master copies of the global variational parameters and priors (device, f64)
  double *W, *vinit, *emit, *prior_tran, *prior_init, *prior_emit;
  double *omega, *omega_prior, *lw;   // mixtures: Dirichlet weights (KE), their prior, E[ln pi] (KE)
  int have_mix;
  double* ada_G; int adagrad;         // AdaGrad-like transition step (hmmsgd_metaobs.py:1036-1040)
  int user_init, have_globals, have_prior;
  // derived per-global-step constants
  float *Pt, *PtT, *pi0;          // exp(E[log A]) row-major, its transpose, exp(E[log pi])
  double *lu, *rowsum, *ckc;      // scratch of the stationary solve / per-(k,d) constants
  double *Rs, *gk, *ck;           // emission constants (see global.cuh)
  double *par2, *ckp;             // diagonal emission constants in the fused kernel's form
  // resident series
  const void* obs; const uint8_t* mask; int obs_dtype; int64_t T_full;
  void* obs_own; uint8_t* mask_own;
  // host-streamed series
  const void* hobs; const uint8_t* hmask; const void* hobs_dev; const uint8_t* hmask_dev;
  int h_dtype; int64_t hT_full; int h_reg_obs, h_reg_mask;
  void* stage_obs; uint8_t* stage_mask; int64_t* stage_src; int64_t* stage_starts;
  double* stage_stats; size_t stage_rows, stage_B;
  void* pin_obs; uint8_t* pin_mask; size_t pin_rows;
  // ring of staging slots for the streamed step (svihmm_prefetch_windows / svihmm_estep_streamed):
  // a slot holds the gathered windows of one minibatch; upcoming minibatches are gathered on
  // `cstream` (back to back, so the host link stays busy) while the current one is processed
  void* sg_obs[SVIHMM_NSLOT]; uint8_t* sg_mask[SVIHMM_NSLOT]; int64_t* sg_src[SVIHMM_NSLOT]; int64_t* sg_dense[SVIHMM_NSLOT];
  size_t sg_rows[SVIHMM_NSLOT], sg_B[SVIHMM_NSLOT];
  int64_t* sg_pin_starts[SVIHMM_NSLOT];               // pinned copies of the window starts
  void* sg_pin_obs[SVIHMM_NSLOT]; uint8_t* sg_pin_mask[SVIHMM_NSLOT]; size_t sg_pin_rows[SVIHMM_NSLOT];   // CPU-gather staging (series that cannot be page-locked: read-only memmaps larger than host memory)
  cudaStream_t cstream, dstream;                      // gather / result read-back streams
  cudaEvent_t ev_gathered[SVIHMM_NSLOT], ev_consumed[SVIHMM_NSLOT], ev_stats, ev_read;
  int sg_valid[SVIHMM_NSLOT], sg_T[SVIHMM_NSLOT], sg_nB[SVIHMM_NSLOT], sg_init, sg_ring;
  int64_t sg_age[SVIHMM_NSLOT], sg_clock;             // last use of each slot (LRU replacement)
  int sg_pending[SVIHMM_NSLOT];                       // staged, not consumed by an E-step yet
  double* pin_stats;
  // peer exchange (svihmm_comm_attach): [16 u64 flags | stats parity 0 | stats parity 1] per rank
  int comm_world, comm_rank; unsigned long long comm_seq; void* comm_peer[8];
  // workspaces
  size_t cap_rows, cap_B, cap_part;
  double *ll_ws, *mx_ws, *seq_ws, *lt_ws;
  double* ell_ws; float *resp_ws, *wq_ws, *part2_ws; size_t cap_rows_mix, cap_part2;   // mixture workspaces
  float *qin_ws, *respin_ws; int64_t* starts_in; size_t cap_qin, cap_respin, cap_startsin;   // buffered windows
  int* e_ws;
  float *dn_b, *dn_a, *dn_r; int* dn_e; uint16_t *dn_q16, *dn_fhi, *dn_flo; size_t cap_dn, cap_dnf;   // dense (tcgen05) recursion tables, tile layout
  float *b_ws, *alpha_ws, *q_ws, *r_ws, *part_ws, *hostq_ws;
  uint8_t* etc_blob; size_t cap_etc;                      // sliced emission factors for k_emit_tc (emit_tc.cuh)
  float *beta_ws, *sb_ws; size_t cap_beta; int last_beta;   // KEEP_LOCALS: normalised backward messages + scale factors
  size_t hostq_cap;
  float *scan_ops, *scan_bound; size_t cap_scan; int scan_min_T; int no_hostreg;
  double* acc_stats[2];                                   // svihmm_svi_run: double-buffered E-step accumulators
  int pdl;                                                // launches carry the programmatic-dependent-launch attribute
  int* status_dev;                                        // device flags raised by the global step (svihmm_check)          // block-parallel scan for long chains (scan16.cuh)
  // batched tensor-core path for K <= 16 (batch16.cuh): (rows, 16) float tables, exponents, row maxima
  float *b16_b, *b16_a, *b16_c; int* b16_E; double* b16_mx; size_t cap_b16; int b16_min_B;
  int last_B, last_T, last_fused;
  int max_smem_optin, fused_attr_set;
  int64_t launches;
  // optional per-phase event timing
  int profiling;
  std::vector<cudaEvent_t>* ev_pool;                 // all events ever created (reused after a read)
  std::vector<int>* ev_phase;                        // phase of each recorded (start, stop) pair
  size_t ev_used;
};

enum { PH_EMIT = 0, PH_FORWARD, PH_BACKWARD, PH_STATS, PH_UPDATE, PH_GATHER, PH_FUSED, PH_OTHER };

__device__ __forceinline__ double ld_obs(const void* obs, int dtype, int64_t idx) {
  return dtype == SVIHMM_F32 ? (double)__ldg((const float*)obs + idx) : __ldg((const double*)obs + idx);
}

// psi(x), float64.  Recurrence up to x >= 10, then the asymptotic series through B14
// (truncation error < 5e-17 there); reflection for x <= 0.
__device__ inline double digamma_d(double x) {
  double r = 0.0;
  if (x <= 0.0) {
    if (x == floor(x)) return nan("");
    r = -M_PI / tan(M_PI * x);
    x = 1.0 - x;
  }
  while (x < 10.0) { r -= 1.0 / x; x += 1.0; }
  const double xi = 1.0 / x, x2 = xi * xi;
  r += log(x) - 0.5 * xi
     - x2 * (1.0 / 12 - x2 * (1.0 / 120 - x2 * (1.0 / 252 - x2 * (1.0 / 240
     - x2 * (1.0 / 132 - x2 * (691.0 / 32760 - x2 * (1.0 / 12)))))));
  return r;
}
