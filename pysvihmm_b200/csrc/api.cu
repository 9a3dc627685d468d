// libsvihmm.so: C ABI (include/svihmm.h) over the sm_100a kernels.  No CPU fallback anywhere:
// every compute entry point enqueues CUDA kernels and fails if there is no device.
#include <string>
#include <string.h>
#include <stdlib.h>
#include <stdarg.h>
#include <algorithm>

#include "common.cuh"
#include "emit.cuh"
#include "fb.cuh"
#include "stats.cuh"
#include "global.cuh"
#include "fused.cuh"
#include "fused_pipe.cuh"
#include "wide64.cuh"
#include "ffbs.cuh"
#include "dense.cuh"
#include "dense_cluster.cuh"
#include "batch16.cuh"
#include "stats_tc.cuh"
#include "emit_tc.cuh"
#include "gth_cluster.cuh"
#include "emit_dense.cuh"
#include "scan16.cuh"
#include "bound.cuh"
#include <cudaTypedefs.h>

static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  g_err = buf;
  return code;
}
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
  return fail(SVIHMM_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define LAUNCHED(ctx) do { (ctx)->launches++; CU(cudaGetLastError()); } while (0)

extern "C" const char* svihmm_last_error(void) { return g_err.c_str(); }
extern "C" int svihmm_version(void) { return 100; }

// ---- per-phase event timing ------------------------------------------------------------------
struct PhaseTimer {
  svihmm_ctx* c; cudaStream_t st; cudaEvent_t stop; bool on;
  PhaseTimer(svihmm_ctx* c_, int phase, cudaStream_t st_) : c(c_), st(st_), stop(nullptr), on(false) {
    if (!c->profiling) return;
    auto& pool = *c->ev_pool;
    while (pool.size() < c->ev_used + 2) {
      cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return; pool.push_back(e);
    }
    cudaEventRecord(pool[c->ev_used], st);
    stop = pool[c->ev_used + 1];
    c->ev_used += 2;
    c->ev_phase->push_back(phase);
    on = true;
  }
  ~PhaseTimer() { if (on) cudaEventRecord(stop, st); }
};

extern "C" int svihmm_set_profiling(svihmm_ctx* c, int on) {
  if (!c) return fail(SVIHMM_EINVAL, "ctx is NULL");
  c->profiling = on ? 1 : 0;
  return SVIHMM_OK;
}

extern "C" const char* svihmm_phase_name(int p) {
  static const char* names[SVIHMM_N_PHASES] = {"emit", "forward", "backward", "stats", "update",
                                               "gather", "fused", "other"};
  return (p >= 0 && p < SVIHMM_N_PHASES) ? names[p] : "?";
}

extern "C" int svihmm_get_phase_ms(svihmm_ctx* c, double* ms, int64_t* counts) {
  if (!c || !ms || !counts) return fail(SVIHMM_EINVAL, "NULL argument");
  for (int i = 0; i < SVIHMM_N_PHASES; ++i) { ms[i] = 0.0; counts[i] = 0; }
  CU(cudaSetDevice(c->device));
  for (size_t i = 0; i < c->ev_phase->size(); ++i) {
    cudaEvent_t a = (*c->ev_pool)[2 * i], b = (*c->ev_pool)[2 * i + 1];
    CU(cudaEventSynchronize(b));
    float t = 0.f;
    CU(cudaEventElapsedTime(&t, a, b));
    ms[(*c->ev_phase)[i]] += t; counts[(*c->ev_phase)[i]]++;
  }
  c->ev_phase->clear(); c->ev_used = 0;
  return SVIHMM_OK;
}

static int next_pow2(int k) { int p = 2; while (p < k) p <<= 1; return p; }

template <typename Tp> static cudaError_t dalloc(Tp** p, size_t n) {
  return cudaMalloc((void**)p, (n ? n : 1) * sizeof(Tp));
}

static int create_impl(svihmm_ctx** out, int device, int K, int D, int kind, int C) {
  if (!out) return fail(SVIHMM_EINVAL, "out is NULL");
  if (C < 1 || C > 64) return fail(SVIHMM_EINVAL, "mixture components C = %d must be in 1..64", C);
  if (C > 1 && kind == SVIHMM_EMIT_CATEGORICAL) return fail(SVIHMM_EUNSUPPORTED, "mixtures of categorical emissions");
  if (K < 1 || D < 1) return fail(SVIHMM_EINVAL, "K (%d) and D (%d) must be >= 1", K, D);
  if (kind != SVIHMM_EMIT_NIW_FULL && kind != SVIHMM_EMIT_NIW_DIAG && kind != SVIHMM_EMIT_CATEGORICAL)
    return fail(SVIHMM_EINVAL, "unknown emission kind %d", kind);
  if (K > 1024) return fail(SVIHMM_EUNSUPPORTED, "K = %d > 1024 (one thread per state in the wide recursion kernels)", K);
  if (kind == SVIHMM_EMIT_NIW_FULL && D > 96) return fail(SVIHMM_EUNSUPPORTED, "full-covariance D = %d > 96", D);
  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(SVIHMM_EINVAL, "device %d out of range (%d devices)", device, ndev);
  CU(cudaSetDevice(device));
  svihmm_ctx* c = (svihmm_ctx*)calloc(1, sizeof(svihmm_ctx));
  if (!c) return fail(SVIHMM_ENOMEM, "calloc");
  CU(cudaDeviceGetAttribute(&c->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
  c->ev_pool = new std::vector<cudaEvent_t>(); c->ev_phase = new std::vector<int>();
  c->device = device; c->K = K; c->D = D; c->kind = kind; c->KP = next_pow2(K);
  c->C = C; c->KE = K * C;
  c->scan_min_T = 4096;
  c->b16_min_B = 4096;     // measured crossover with the one-CTA-per-window kernel at K16/D8/T512
  const size_t KE = (size_t)c->KE;
  c->DD = kind == SVIHMM_EMIT_NIW_FULL ? D * D : (kind == SVIHMM_EMIT_NIW_DIAG ? D : 0);
  c->OD = kind == SVIHMM_EMIT_CATEGORICAL ? 1 : D;
  c->plen = kind == SVIHMM_EMIT_NIW_FULL ? (size_t)D + (size_t)D * D + 2
          : (kind == SVIHMM_EMIT_NIW_DIAG ? (size_t)4 * D : (size_t)D);
  c->nfeat = K + 1 + D + c->DD;
  c->slen = (size_t)K * K + KE + KE * D + KE * c->DD + K + 4;
  const size_t KK = (size_t)K * K;
  const size_t rs = kind == SVIHMM_EMIT_NIW_FULL ? KE * D * (D + 1) / 2 : KE * D;
  CU(dalloc(&c->W, KK)); CU(dalloc(&c->vinit, 2 * (size_t)K)); CU(dalloc(&c->emit, KE * c->plen));
  CU(dalloc(&c->prior_tran, KK)); CU(dalloc(&c->prior_init, (size_t)K)); CU(dalloc(&c->prior_emit, KE * c->plen));
  CU(dalloc(&c->omega, KE)); CU(dalloc(&c->omega_prior, KE)); CU(dalloc(&c->lw, KE));
  CU(dalloc(&c->ada_G, KK));
  CU(dalloc(&c->Pt, KK)); CU(dalloc(&c->PtT, KK)); CU(dalloc(&c->pi0, (size_t)K));
  CU(dalloc(&c->lu, 2 * (size_t)K * (K + 1) + 8 * (size_t)K + 16)); CU(dalloc(&c->rowsum, (size_t)K)); CU(dalloc(&c->ckc, 2 * KE * D));
  CU(dalloc(&c->par2, 2 * KE * D)); CU(dalloc(&c->ckp, KE));
  CU(dalloc(&c->Rs, rs)); CU(dalloc(&c->gk, KE * D)); CU(dalloc(&c->ck, KE));
  CU(dalloc(&c->stage_stats, c->slen));
  CU(dalloc(&c->status_dev, (size_t)1)); CU(cudaMemset(c->status_dev, 0, sizeof(int)));
  *out = c;
  return SVIHMM_OK;
}

extern "C" int svihmm_create(svihmm_ctx** out, int device, int K, int D, int kind) {
  return create_impl(out, device, K, D, kind, 1);
}

extern "C" int svihmm_create_mix(svihmm_ctx** out, int device, int K, int D, int kind, int C) {
  return create_impl(out, device, K, D, kind, C);
}

static void sg_teardown(svihmm_ctx* c);

static void free_streamed(svihmm_ctx* c) {
  if (c->h_reg_obs) cudaHostUnregister((void*)c->hobs);
  if (c->h_reg_mask) cudaHostUnregister((void*)c->hmask);
  c->h_reg_obs = c->h_reg_mask = 0; c->hobs = nullptr; c->hmask = nullptr;
  c->hobs_dev = nullptr; c->hmask_dev = nullptr;
}

extern "C" int svihmm_destroy(svihmm_ctx* c) {
  if (!c) return SVIHMM_OK;
  cudaSetDevice(c->device);
  sg_teardown(c);
  free_streamed(c);
  void* ptrs[] = {c->W, c->vinit, c->emit, c->prior_tran, c->prior_init, c->prior_emit, c->Pt, c->PtT,
                  c->pi0, c->lu, c->rowsum, c->ckc, c->par2, c->ckp, c->Rs, c->gk, c->ck, c->obs_own, c->mask_own, c->stage_obs,
                  c->stage_mask, c->stage_src, c->stage_starts, c->stage_stats, c->ll_ws, c->mx_ws, c->lt_ws, c->e_ws,
                  c->seq_ws, c->b_ws, c->alpha_ws, c->q_ws, c->r_ws, c->part_ws, c->hostq_ws,
                  c->omega, c->omega_prior, c->lw, c->ell_ws, c->resp_ws, c->wq_ws, c->part2_ws,
                  c->qin_ws, c->respin_ws, c->starts_in, c->ada_G, c->beta_ws, c->sb_ws,
                  c->b16_b, c->b16_a, c->b16_c, c->b16_E, c->b16_mx, c->scan_ops, c->scan_bound, c->status_dev, c->acc_stats[0], c->acc_stats[1], c->dn_b, c->dn_a, c->dn_r, c->dn_e, c->dn_q16, c->dn_fhi, c->dn_flo, c->etc_blob};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (c->pin_obs) cudaFreeHost(c->pin_obs);
  if (c->pin_mask) cudaFreeHost(c->pin_mask);
  for (cudaEvent_t e : *c->ev_pool) cudaEventDestroy(e);
  delete c->ev_pool; delete c->ev_phase;
  free(c);
  return SVIHMM_OK;
}

extern "C" size_t svihmm_emit_param_len(const svihmm_ctx* c) { return c ? c->plen : 0; }
extern "C" size_t svihmm_stats_len(const svihmm_ctx* c) { return c ? c->slen : 0; }
extern "C" int64_t svihmm_launch_count(const svihmm_ctx* c) { return c ? c->launches : 0; }

static size_t esize(int dtype) { return dtype == SVIHMM_F32 ? 4 : 8; }

extern "C" int svihmm_set_series(svihmm_ctx* c, const void* obs, int64_t T_full, int dtype,
                                 const uint8_t* mask, int loc, void* stream) {
  if (!c || !obs) return fail(SVIHMM_EINVAL, "ctx/obs is NULL");
  if (T_full < 1) return fail(SVIHMM_EINVAL, "T_full = %lld", (long long)T_full);
  if (dtype != SVIHMM_F32 && dtype != SVIHMM_F64) return fail(SVIHMM_EINVAL, "dtype %d", dtype);
  CU(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  if (c->obs_own) { CU(cudaFree(c->obs_own)); c->obs_own = nullptr; }
  if (c->mask_own) { CU(cudaFree(c->mask_own)); c->mask_own = nullptr; }
  if (loc == SVIHMM_LOC_DEVICE) {
    c->obs = obs; c->mask = mask;
  } else {
    const size_t nb = (size_t)T_full * c->OD * esize(dtype);
    CU(cudaMalloc(&c->obs_own, nb));
    CU(cudaMemcpyAsync(c->obs_own, obs, nb, cudaMemcpyHostToDevice, st));
    c->obs = c->obs_own; c->mask = nullptr;
    if (mask) {
      CU(cudaMalloc((void**)&c->mask_own, (size_t)T_full));
      CU(cudaMemcpyAsync(c->mask_own, mask, (size_t)T_full, cudaMemcpyHostToDevice, st));
      c->mask = c->mask_own;
    }
    CU(cudaStreamSynchronize(st));
  }
  c->obs_dtype = dtype; c->T_full = T_full;
  return SVIHMM_OK;
}

extern "C" int svihmm_set_series_streamed(svihmm_ctx* c, const void* obs_host, int64_t T_full, int dtype,
                                          const uint8_t* mask_host) {
  if (!c || !obs_host) return fail(SVIHMM_EINVAL, "ctx/obs is NULL");
  if (dtype != SVIHMM_F32 && dtype != SVIHMM_F64) return fail(SVIHMM_EINVAL, "dtype %d", dtype);
  CU(cudaSetDevice(c->device));
  free_streamed(c);
  c->hobs = obs_host; c->hmask = mask_host; c->h_dtype = dtype; c->hT_full = T_full;
  // Page-lock + map the caller's buffer so the GPU gathers each step's windows itself.  If the
  // registration is refused (e.g. read-only mapping) the CPU-gather + pinned-staging path is used.
  const size_t nb = (size_t)T_full * c->OD * esize(dtype);
  // (a read-only mapping, e.g. np.memmap(mode='r') of gen_synthetic.read_data_mmap, needs the ReadOnly flag)
  if (c->no_hostreg) return SVIHMM_OK;        // tuning: CPU gather + pinned staging (series larger than host memory)
  if (cudaHostRegister((void*)obs_host, nb, cudaHostRegisterMapped) == cudaSuccess ||
      ((void)cudaGetLastError(), cudaHostRegister((void*)obs_host, nb, cudaHostRegisterMapped | cudaHostRegisterReadOnly) == cudaSuccess)) {
    c->h_reg_obs = 1;
    void* dp = nullptr;
    if (cudaHostGetDevicePointer(&dp, (void*)obs_host, 0) == cudaSuccess) c->hobs_dev = dp;
  }
  (void)cudaGetLastError();
  if (mask_host && c->hobs_dev) {
    if (cudaHostRegister((void*)mask_host, (size_t)T_full, cudaHostRegisterMapped) == cudaSuccess ||
        ((void)cudaGetLastError(), cudaHostRegister((void*)mask_host, (size_t)T_full, cudaHostRegisterMapped | cudaHostRegisterReadOnly) == cudaSuccess)) {
      c->h_reg_mask = 1;
      void* dp = nullptr;
      if (cudaHostGetDevicePointer(&dp, (void*)mask_host, 0) == cudaSuccess) c->hmask_dev = (const uint8_t*)dp;
    }
    (void)cudaGetLastError();
    if (!c->hmask_dev) c->hobs_dev = nullptr;   // all-or-nothing: fall back to the CPU gather
  }
  return SVIHMM_OK;
}

static int copy_in(void* dst, const void* src, size_t nb, int loc, cudaStream_t st) {
  CU(cudaMemcpyAsync(dst, src, nb, loc == SVIHMM_LOC_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
  if (loc == SVIHMM_LOC_HOST) CU(cudaStreamSynchronize(st));   // caller may free its buffer on return
  return SVIHMM_OK;
}

extern "C" int svihmm_set_prior(svihmm_ctx* c, const double* prior_tran, const double* prior_init,
                                const double* prior_emit, int loc, void* stream) {
  if (!c || !prior_tran || !prior_emit) return fail(SVIHMM_EINVAL, "NULL argument");
  CU(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if ((rc = copy_in(c->prior_tran, prior_tran, sizeof(double) * c->K * c->K, loc, st))) return rc;
  if ((rc = copy_in(c->prior_emit, prior_emit, sizeof(double) * c->KE * c->plen, loc, st))) return rc;
  if (prior_init) { if ((rc = copy_in(c->prior_init, prior_init, sizeof(double) * c->K, loc, st))) return rc; }
  else {
    double* ones = (double*)malloc(sizeof(double) * c->K);
    for (int i = 0; i < c->K; ++i) ones[i] = 1.0;
    rc = copy_in(c->prior_init, ones, sizeof(double) * c->K, SVIHMM_LOC_HOST, st);
    free(ones);
    if (rc) return rc;
  }
  c->have_prior = 1;
  return SVIHMM_OK;
}

// One launch: (optional) global update of the master parameters + all derived per-step constants.
// blocks of k_global_step (= flags per source rank in the exchange area): transitions + emission
// blocks + mixture weights, the same on every rank
static int comm_blocks(const svihmm_ctx* c) {
  const int KE = c->KE;
  const int nblk = c->kind == SVIHMM_EMIT_NIW_DIAG ? std::max(1, std::min(KE, (KE * c->D + 255) / 256)) : KE;
  return 1 + nblk + (c->C > 1 ? 1 : 0);
}
static int run_global(svihmm_ctx* c, int mode, const double* stats, double lrate, double bA, double bE,
                      cudaStream_t st, bool peers = false, double* zero_buf = nullptr) {
  const int K = c->K, D = c->D;
  GlobalArgs ga;
  ga.K = K; ga.D = D; ga.DD = c->DD; ga.diag = c->kind == SVIHMM_EMIT_NIW_DIAG; ga.cat = c->kind == SVIHMM_EMIT_CATEGORICAL; ga.mode = mode;
  ga.user_init = c->user_init; ga.plen = c->plen;
  ga.KE = c->KE; ga.C = c->C; ga.omega = c->omega; ga.omega_prior = c->omega_prior; ga.lw = c->lw;
  ga.ada_G = c->adagrad ? c->ada_G : nullptr;
  ga.world = 1; ga.rank = 0; ga.nb = 0; ga.seq = 0; ga.red_out = nullptr; ga.stats_local = nullptr; ga.slen = c->slen;
  ga.W = c->W; ga.vinit = c->vinit; ga.emit = c->emit;
  ga.prior_tran = c->prior_tran; ga.prior_init = c->prior_init; ga.prior_emit = c->prior_emit;
  ga.stats = stats ? stats : c->stage_stats;      // unused in GM_PREP
  ga.lrate = lrate; ga.bA = bA; ga.bE = bE;
  ga.gth = c->lu; ga.rowsum = c->rowsum; ga.ckc = c->ckc;
  ga.status = c->status_dev; ga.zero_buf = zero_buf;
  static const bool no_gc = getenv("SVIHMM_NO_GTH_CLUSTER") != nullptr;   // A/B switch, read once
  ga.gth_ext = (!no_gc && !ga.user_init && K > 64 && K <= GC_KMAX) ? 1 : 0;   // at K = 64 the one-CTA elimination is as fast (c3: 0.146 vs 0.171 ms)
  ga.Pt = c->Pt; ga.PtT = c->PtT; ga.pi0 = c->pi0; ga.Rs = c->Rs; ga.gk = c->gk; ga.ck = c->ck; ga.par2 = c->par2; ga.ckp = c->ckp;
  const int KE = c->KE;
  if (c->C > 1 && mode != GM_PREP && mode != GM_SVI) return fail(SVIHMM_EUNSUPPORTED, "mixture emissions support the SVI update only");
  if (c->C > 1 && !c->have_mix) return fail(SVIHMM_ESTATE, "svihmm_set_mix_weights has not been called");
  const int nblk = ga.diag ? std::max(1, std::min(KE, (KE * D + 255) / 256)) : KE;   // full / categorical: one block per component
  size_t smem = (2 * (size_t)K + 2) * sizeof(double);
  if (!ga.diag && !ga.cat) smem = std::max(smem, (2 * (size_t)D * D + 3 * (size_t)D) * sizeof(double));
  if (K > 32 && K <= GTH_SMEM_KMAX) smem = std::max(smem, (2 * (size_t)K + 2 + (size_t)K * K) * sizeof(double));   // elimination matrix of block 0
  if (smem > 48 * 1024) CU(cudaFuncSetAttribute(k_global_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (peers) {
    ga.world = c->comm_world; ga.rank = c->comm_rank; ga.seq = ++c->comm_seq; ga.red_out = c->stage_stats;
    ga.nb = comm_blocks(c);
    ga.stats_local = stats; ga.stats = c->stage_stats;          // the update reads the sums from red_out
    for (int p = 0; p < c->comm_world; ++p) ga.xbase[p] = (unsigned long long*)c->comm_peer[p];
  }
  static const bool gdbg = getenv("SVIHMM_GLOBAL_DBG") != nullptr;
  ga.dbg = nullptr;
  if (gdbg) CU(cudaMalloc((void**)&ga.dbg, 128));
  {
    PhaseTimer pt(c, PH_UPDATE, st);
    // K*K digamma threads + one warp for the stationary vector
    // K*K digamma threads + one warp for the stationary vector
    const int nthr = std::min(512, ((K * K + 31) / 32) * 32 + 32);
    if (c->pdl) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(1 + nblk + (c->C > 1 ? 1 : 0)); cfg.blockDim = dim3(std::max(nthr, 128));
      cfg.dynamicSmemBytes = smem; cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      CU(cudaLaunchKernelEx(&cfg, k_global_step, ga, nblk));
    } else {
      k_global_step<<<1 + nblk + (c->C > 1 ? 1 : 0), std::max(nthr, 128), smem, st>>>(ga, nblk);
    }
    LAUNCHED(c);
    if (ga.gth_ext) {
      // stationary vector on a cluster of 8 CTAs, matrix in distributed shared memory (gth_cluster.cuh)
      const size_t gsm = gth_cluster_smem(K);
      static bool gc_attr = false;
      if (!gc_attr) { CU(cudaFuncSetAttribute(k_gth_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gth_cluster_smem(GC_KMAX))); gc_attr = true; }
      k_gth_cluster<<<GC_CTAS, GC_NT, gsm, st>>>(K, c->lu, c->W, c->rowsum, c->Pt, c->PtT, c->vinit, c->pi0);
      LAUNCHED(c);
    }
  }
  if (gdbg) {
    long long h[16];
    CU(cudaStreamSynchronize(st));
    CU(cudaMemcpy(h, ga.dbg, 128, cudaMemcpyDeviceToHost));
    fprintf(stderr, "[global dbg] stationary warp done %lld cycles after the row sums; into pi0_section=%lld n2 loop=%lld shuffles=%lld sqrt=%lld\n",
            h[14] - h[1], h[11] - h[2], h[12] - h[11], h[13] - h[12], h[6] - h[13]);
    fprintf(stderr, "[global dbg] squarings=%lld pi0 section: norm2=%lld store=%lld n1=%lld dgs=%lld rest=%lld\n", h[10], h[6] - h[2], h[7] - h[6], h[8] - h[7], h[9] - h[8], h[3] - h[9]);
    CU(cudaFree(ga.dbg));
    fprintf(stderr, "[global dbg] mode=%d cycles: update+rowsum=%lld P||stationary=%lld gth=%lld pi0=%lld | emission block=%lld\n", mode,
            h[1] - h[0], h[2] - h[1], 0LL, h[3] - h[2], h[5] - h[4]);
  }
  return SVIHMM_OK;
}

extern "C" int svihmm_set_globals(svihmm_ctx* c, const double* var_tran, const double* var_init,
                                  const double* emit, int loc, void* stream) {
  if (!c || !var_tran || !emit) return fail(SVIHMM_EINVAL, "NULL argument");
  CU(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if ((rc = copy_in(c->W, var_tran, sizeof(double) * c->K * c->K, loc, st))) return rc;
  if ((rc = copy_in(c->emit, emit, sizeof(double) * c->KE * c->plen, loc, st))) return rc;
  c->user_init = var_init != nullptr;
  if (var_init && (rc = copy_in(c->vinit + c->K, var_init, sizeof(double) * c->K, loc, st))) return rc;
  c->have_globals = 1;
  if (c->C > 1 && !c->have_mix) return SVIHMM_OK;      // constants are derived once the mixture weights arrive
  return run_global(c, GM_PREP, nullptr, 0.0, 0.0, 0.0, st);
}

extern "C" int svihmm_set_mix_weights(svihmm_ctx* c, const double* omega, const double* omega_prior,
                                      int loc, void* stream) {
  if (!c || !omega) return fail(SVIHMM_EINVAL, "NULL argument");
  if (c->C < 2) return fail(SVIHMM_ESTATE, "the context was not created with svihmm_create_mix (C >= 2)");
  CU(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if ((rc = copy_in(c->omega, omega, sizeof(double) * c->KE, loc, st))) return rc;
  if (omega_prior && (rc = copy_in(c->omega_prior, omega_prior, sizeof(double) * c->KE, loc, st))) return rc;
  if (!omega_prior && !c->have_mix) return fail(SVIHMM_EINVAL, "omega_prior is required on the first call");
  c->have_mix = 1;
  return c->have_globals ? run_global(c, GM_PREP, nullptr, 0.0, 0.0, 0.0, st) : SVIHMM_OK;
}

extern "C" int svihmm_get_mix_weights(svihmm_ctx* c, double* omega, int loc, void* stream) {
  if (!c || !omega) return fail(SVIHMM_EINVAL, "NULL argument");
  if (c->C < 2 || !c->have_mix) return fail(SVIHMM_ESTATE, "no mixture weights set");
  CU(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaMemcpyAsync(omega, c->omega, sizeof(double) * c->KE,
                     loc == SVIHMM_LOC_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
  if (loc == SVIHMM_LOC_HOST) CU(cudaStreamSynchronize(st));
  return SVIHMM_OK;
}

extern "C" int svihmm_get_globals(svihmm_ctx* c, double* var_tran, double* var_init, double* emit,
                                  int loc, void* stream) {
  if (!c) return fail(SVIHMM_EINVAL, "ctx is NULL");
  if (!c->have_globals) return fail(SVIHMM_ESTATE, "svihmm_set_globals has not been called");
  CU(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const cudaMemcpyKind kd = loc == SVIHMM_LOC_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  if (var_tran) CU(cudaMemcpyAsync(var_tran, c->W, sizeof(double) * c->K * c->K, kd, st));
  if (var_init) CU(cudaMemcpyAsync(var_init, c->vinit, sizeof(double) * c->K, kd, st));
  if (emit) CU(cudaMemcpyAsync(emit, c->emit, sizeof(double) * c->KE * c->plen, kd, st));
  if (loc == SVIHMM_LOC_HOST) CU(cudaStreamSynchronize(st));
  return SVIHMM_OK;
}

static int ensure_ws(svihmm_ctx* c, int B, int T, bool need_r) {
  const size_t rows = (size_t)B * T, K = c->K;
  if (rows > c->cap_rows) {
    void* olds[] = {c->ll_ws, c->mx_ws, c->b_ws, c->alpha_ws, c->q_ws, c->r_ws, c->lt_ws, c->e_ws};
    for (void* p : olds) if (p) CU(cudaFree(p));
    c->ll_ws = nullptr; c->mx_ws = nullptr; c->b_ws = nullptr; c->alpha_ws = nullptr; c->q_ws = nullptr; c->r_ws = nullptr;
    c->lt_ws = nullptr; c->e_ws = nullptr;
    c->cap_rows = 0;
    CU(dalloc(&c->ll_ws, rows * K)); CU(dalloc(&c->mx_ws, 2 * rows));
    CU(dalloc(&c->lt_ws, rows)); CU(dalloc(&c->e_ws, rows));
    CU(dalloc(&c->b_ws, rows * K)); CU(dalloc(&c->alpha_ws, rows * K)); CU(dalloc(&c->q_ws, rows * K));
    c->cap_rows = rows;
  }
  if (need_r && !c->r_ws) CU(dalloc(&c->r_ws, c->cap_rows * K));
  if ((size_t)B > c->cap_B) {
    if (c->seq_ws) CU(cudaFree(c->seq_ws));
    c->seq_ws = nullptr; c->cap_B = 0;
    CU(dalloc(&c->seq_ws, 2 * (size_t)B));
    c->cap_B = B;
  }
  return SVIHMM_OK;
}

template <int KP>
static void launch_fb(svihmm_ctx* c, int B, int T, float* q, float* r, cudaStream_t st, float* beta_out, float* sb_out) {
  const int G = 32 / KP;
  const int warps = (B + G - 1) / G;
  // few chains: one warp per CTA so that every chain gets a scheduler of its own
  const int wpb = warps <= 4 * 148 ? 1 : 4;
  const int grid = (warps + wpb - 1) / wpb;
  float* cs = (float*)(c->mx_ws + (size_t)B * T);   // second half of mx_ws holds the scale factors
  { PhaseTimer pt(c, PH_FORWARD, st);
    k_forward<KP><<<grid, wpb * 32, 0, st>>>(B, T, c->K, c->Pt, c->pi0, c->b_ws, c->alpha_ws, cs);
    c->launches++; }
  { PhaseTimer pt(c, PH_BACKWARD, st);
    k_backward<KP><<<grid, wpb * 32, 0, st>>>(B, T, c->K, c->Pt, c->b_ws, c->alpha_ws, q, r, beta_out, sb_out);
    c->launches++; }
}

template <int KP>
static cudaError_t launch_fused(const FusedArgs& fa, size_t smem, cudaStream_t st, bool set_attr) {
  if (set_attr) {
    cudaError_t e = cudaFuncSetAttribute(k_estep_fused<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  k_estep_fused<KP><<<fa.B, FUSED_NT, smem, st>>>(fa);
  return cudaGetLastError();
}

template <int KP, int NTE>
static cudaError_t launch_pipe(const FusedArgs& fa, size_t smem, cudaStream_t st, bool set_attr, bool pdl = false) {
  if (set_attr) {
    cudaError_t e = cudaFuncSetAttribute(k_estep_pipe<KP, NTE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  if (pdl) {                 // svihmm_svi_run: the prologue may overlap the preceding update kernel
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(fa.B); cfg.blockDim = dim3(FP_NT); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k_estep_pipe<KP, NTE>, fa);
  }
  k_estep_pipe<KP, NTE><<<fa.B, FP_NT, smem, st>>>(fa);
  return cudaGetLastError();
}

// Pipelined single-kernel E-step (fused_pipe.cuh): the diagonal model with K <= 32 whose window
// fits in shared memory and whose emission features [x | x^2 | w] fit the tensor-core statistics
// tiles instantiated below (D <= 16, or D <= 8 for 16 < K <= 32); barriers within the budget.
static int pipe_nte(const svihmm_ctx* c) { return 2 * ((c->D + 7) / 8) + 1; }
static bool pipe_eligible(const svihmm_ctx* c, int T, unsigned flags, size_t* smem_out) {
  static const bool off = getenv("SVIHMM_NO_PIPE") != nullptr;
  if (off || c->K > 32 || c->C > 1 || c->kind != SVIHMM_EMIT_NIW_DIAG || (flags & (SVIHMM_EXACT_XI | SVIHMM_KEEP_LOCALS))) return false;
  const int nte = pipe_nte(c);
  if (nte > 5 || (c->K > 16 && nte > 3)) return false;
  const PipeSmem L = pipe_smem_layout(T, c->K, c->D, c->D, 1);
  if (L.total > (size_t)c->max_smem_optin) return false;
  const int nwk = c->K > 16 ? 128 : 192;
  const int npairs = (T + 1) / 2, nrounds = (npairs + nwk / 2 - 1) / (nwk / 2), ntiles = (T - T / 2 + FP_TB - 1) / FP_TB;
  if (nrounds + ntiles > FP_MAXBAR || ntiles > 190) return false;
  *smem_out = L.total;
  return true;
}

// Single-kernel E-step (fused.cuh) when the window fits in shared memory: K <= 32, the three
// T*K float tables + per-row scalars + emission constants <= the opt-in limit.
static bool fused_eligible(const svihmm_ctx* c, int T, unsigned flags, size_t* smem_out) {
  if (c->K > 32 || c->C > 1 || c->kind == SVIHMM_EMIT_CATEGORICAL || (flags & (SVIHMM_EXACT_XI | SVIHMM_KEEP_LOCALS))) return false;
  const int diag = c->kind == SVIHMM_EMIT_NIW_DIAG;
  const int tri = diag ? c->D : c->D * (c->D + 1) / 2;
  const FusedSmem L = fused_smem_layout(T, c->K, c->D, tri, diag);
  if (L.total > (size_t)c->max_smem_optin) return false;
  *smem_out = L.total;
  return true;
}

static int estep_fused(svihmm_ctx* c, const void* obs, int dtype, const uint8_t* mask,
                       const int64_t* starts, int B, int T, float* var_x_out, double* stats_out,
                       unsigned flags, size_t smem, cudaStream_t st, bool pipe) {
  const int K = c->K, D = c->D;
  if ((size_t)B > c->cap_B) {
    if (c->seq_ws) CU(cudaFree(c->seq_ws));
    c->seq_ws = nullptr; c->cap_B = 0;
    CU(dalloc(&c->seq_ws, 2 * (size_t)B));
    c->cap_B = B;
  }
  PhaseTimer pt(c, PH_FUSED, st);
  const bool pdl = pipe && c->pdl;      // svihmm_svi_run: the accumulator was zeroed by the previous update kernel
  if (!pdl) CU(cudaMemsetAsync(stats_out, 0, sizeof(double) * c->slen, st));
  FusedArgs fa;
  fa.B = B; fa.T = T; fa.K = K; fa.D = D; fa.DD = c->DD;
  fa.diag = c->kind == SVIHMM_EMIT_NIW_DIAG;
  fa.wrap = (flags & SVIHMM_WRAP) ? 1 : 0; fa.add_prior = (flags & SVIHMM_ADD_PRIOR) ? 1 : 0;
  fa.mask_ll = (flags & SVIHMM_MASK_LL) ? 1 : 0;
  fa.tri = fa.diag ? D : D * (D + 1) / 2;
  fa.obs = obs; fa.dtype = dtype; fa.mask = mask; fa.starts = starts;
  fa.Pt = c->Pt; fa.pi0 = c->pi0; fa.Rs = fa.diag ? c->par2 : c->Rs; fa.gk = c->gk; fa.ck = fa.diag ? c->ckp : c->ck; fa.prior_tran = c->prior_tran;
  fa.var_x_out = var_x_out; fa.stats_out = stats_out; fa.seq = c->seq_ws;
  fa.o_n = (size_t)K * K; fa.o_sx = fa.o_n + K; fa.o_sxx = fa.o_sx + (size_t)K * D;
  fa.o_q0 = fa.o_sxx + (size_t)K * c->DD; fa.o_tail = fa.o_q0 + K;
  static const bool dbg_on = getenv("SVIHMM_FUSED_DBG") != nullptr;
  fa.dbg = nullptr;
  if (dbg_on) CU(cudaMalloc((void**)&fa.dbg, sizeof(long long) * 16 * B));
  const bool set_attr = smem > 48 * 1024;
  cudaError_t e;
  if (pipe) {
    const bool wide = pipe_nte(c) > 3;                  // 8 < D <= 16
    switch (c->KP) {
      case 2:
      case 4: e = wide ? launch_pipe<4, 5>(fa, smem, st, set_attr, pdl) : launch_pipe<4, 3>(fa, smem, st, set_attr, pdl); break;
      case 8: e = wide ? launch_pipe<8, 5>(fa, smem, st, set_attr, pdl) : launch_pipe<8, 3>(fa, smem, st, set_attr, pdl); break;
      case 16: e = wide ? launch_pipe<16, 5>(fa, smem, st, set_attr, pdl) : launch_pipe<16, 3>(fa, smem, st, set_attr, pdl); break;
      default: e = launch_pipe<32, 3>(fa, smem, st, set_attr, pdl); break;
    }
  } else {
    switch (c->KP) {
      case 2:
      case 4: e = launch_fused<4>(fa, smem, st, set_attr); break;
      case 8: e = launch_fused<8>(fa, smem, st, set_attr); break;
      case 16: e = launch_fused<16>(fa, smem, st, set_attr); break;
      default: e = launch_fused<32>(fa, smem, st, set_attr); break;
    }
  }
  if (e != cudaSuccess) return fail(SVIHMM_ECUDA, "fused E-step launch failed: %s", cudaGetErrorString(e));
  if (dbg_on && pipe) {
    std::vector<long long> h((size_t)16 * B);
    CU(cudaStreamSynchronize(st));
    CU(cudaMemcpy(h.data(), fa.dbg, sizeof(long long) * 16 * B, cudaMemcpyDeviceToHost));
    double m[6] = {0, 0, 0, 0, 0, 0}, q[5] = {0, 0, 0, 0, 0};
    for (int b = 0; b < B; ++b) for (int i = 1; i < 6; ++i) m[i] += (double)(h[16 * b + i] - h[16 * b]) / B;
    for (int b = 0; b < B; ++b) for (int i = 0; i < 5; ++i) q[i] += (double)h[16 * b + 8 + i] / B;
    long long gmin = h[13], gmaxs = h[13], gmaxe = h[14];
    for (int b = 0; b < B; ++b) { gmin = std::min(gmin, h[16 * b + 13]); gmaxs = std::max(gmaxs, h[16 * b + 13]); gmaxe = std::max(gmaxe, h[16 * b + 14]); }
    fprintf(stderr, "[pipe dbg] globaltimer: last CTA start - first CTA start = %.2f us, last CTA end - first CTA start = %.2f us\n",
            (gmaxs - gmin) * 1e-3, (gmaxe - gmin) * 1e-3);
    fprintf(stderr, "[pipe dbg] worker 0 totals over tiles: wait=%.0f C1=%.0f bar=%.0f C2=%.0f C3+prefetch=%.0f\n", q[0], q[1], q[2], q[3], q[4]);
    fprintf(stderr, "[pipe dbg] B=%d T=%d smem=%zu cycles since CTA start: chain start=%.0f chain end=%.0f (%.1f/step) | workers: phase A done=%.0f last tile done=%.0f | before final atomics=%.0f\n",
            B, T, smem, m[1], m[2], (m[2] - m[1]) / (T > 1 ? T - 1 : 1), m[3], m[4], m[5]);
  }
  if (dbg_on && !pipe) {   // debug: mean clock cycles per phase over the CTAs of this launch
    std::vector<long long> h((size_t)8 * B);
    CU(cudaStreamSynchronize(st));
    CU(cudaMemcpy(h.data(), fa.dbg, sizeof(long long) * 8 * B, cudaMemcpyDeviceToHost));
    double ph[5] = {0, 0, 0, 0, 0};
    for (int b = 0; b < B; ++b) for (int i = 0; i < 5; ++i) ph[i] += (double)(h[8 * b + i + 1] - h[8 * b + i]) / B;
    fprintf(stderr, "[fused dbg] B=%d T=%d smem=%zu cycles: A(emit)=%.0f B(chains)=%.0f C1(q,out,logZ)=%.0f C2(tran stat)=%.0f C3(emit stats)=%.0f\n",
            B, T, smem, ph[0], ph[1], ph[2], ph[3], ph[4]);
  }
  if (dbg_on) CU(cudaFree(fa.dbg));
  c->launches++;
  c->last_B = B; c->last_T = T; c->last_fused = 1;
  return SVIHMM_OK;
}

// Buffered meta-observations (growBuffer, hmmsgd_metaobs.py:932-1008): the recursions run on
// windows of T rows but only the inner T - 2*trim rows feed the statistics.  The inner rows of a
// (B, T, W) table are compacted into a dense (B, T - 2*trim, W) one and the window starts shifted,
// so that the statistics kernels see an ordinary minibatch.
__global__ void __launch_bounds__(256)
k_trim_rows(int B, int T, int trim, int W, const float* __restrict__ src, float* __restrict__ dst) {
  const int Ts = T - 2 * trim;
  const int64_t n = (int64_t)B * Ts * W;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = e / W; const int j = (int)(e - row * W);
    const int b = (int)(row / Ts); const int t = (int)(row - (int64_t)b * Ts);
    dst[e] = src[((int64_t)b * T + trim + t) * W + j];
  }
}
__global__ void k_shift_starts(int B, int trim, const int64_t* __restrict__ src, int64_t* __restrict__ dst) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) dst[b] = src[b] + trim;
}

static int trim_table(svihmm_ctx* c, int B, int T, int trim, int W, const float* src, float** ws, size_t* cap,
                      cudaStream_t st) {
  const size_t need = (size_t)B * (T - 2 * trim) * W;
  if (need > *cap) {
    if (*ws) CU(cudaFree(*ws));
    *ws = nullptr; *cap = 0;
    CU(dalloc(ws, need));
    *cap = need;
  }
  k_trim_rows<<<(unsigned)std::min<size_t>((need + 255) / 256, 148 * 16), 256, 0, st>>>(B, T, trim, W, src, *ws);
  LAUNCHED(c);
  return SVIHMM_OK;
}

// launch k_stats for one column range with row splits sized for ~4 CTAs per SM; returns nsplit
static int launch_kstats(svihmm_ctx* c, StatsArgs a, int Kleft, int n_lo, int n_hi, float** part, size_t* cap,
                         int64_t* nsplit_out, cudaStream_t st) {
  const int TM = Kleft <= 16 ? 16 : (Kleft <= 32 ? 32 : 64);
  const int tiles_m = (Kleft + TM - 1) / TM, tiles_n = (n_hi - n_lo + ST_TN - 1) / ST_TN;
  const int64_t chunks = (a.R + ST_RC - 1) / ST_RC;
  int64_t nsplit = (4 * 148 + (int64_t)tiles_m * tiles_n - 1) / ((int64_t)tiles_m * tiles_n);
  if (nsplit > chunks) nsplit = chunks;
  if (nsplit < 1) nsplit = 1;
  const int64_t rps = ((chunks + nsplit - 1) / nsplit) * ST_RC;
  nsplit = (a.R + rps - 1) / rps;
  a.rows_per_split = rps; a.n_lo = n_lo; a.n_hi = n_hi;
  const size_t need = (size_t)nsplit * Kleft * a.N;
  if (need > *cap) {
    if (*part) CU(cudaFree(*part));
    *part = nullptr; *cap = 0;
    CU(dalloc(part, need));
    *cap = need;
  }
  a.part = *part;
  dim3 grid(tiles_n, tiles_m, (unsigned)nsplit);
  const size_t xsm = (size_t)ST_RC * a.D * sizeof(float);
  if (TM == 16) k_stats<16><<<grid, 256, xsm, st>>>(a);
  else if (TM == 32) k_stats<32><<<grid, 256, xsm, st>>>(a);
  else k_stats<64><<<grid, 256, xsm, st>>>(a);
  LAUNCHED(c);
  *nsplit_out = nsplit;
  return SVIHMM_OK;
}

// Mixture statistics: transitions from q, NIW statistics of the K*C components weighted by
// q[t,k] * resp[t,k,c] (util.py:73-83 with the responsibilities of labels.py:52-65).
static int stats_mix(svihmm_ctx* c, const void* obs, int dtype, const uint8_t* mask, const int64_t* starts,
                     int B, int T, const float* q, int Tfull, int trim, double* stats_out, unsigned flags,
                     cudaStream_t st) {
  const int K = c->K, KE = c->KE, D = c->D;
  const int64_t R = (int64_t)B * T;
  const float* resp = c->resp_ws;
  if (trim > 0) {                                   // responsibilities of the inner rows only
    int rc_ = trim_table(c, B, Tfull, trim, KE, c->resp_ws, &c->respin_ws, &c->cap_respin, st);
    if (rc_) return rc_;
    resp = c->respin_ws;
  }
  k_mix_weights<<<(unsigned)((R * KE + 255) / 256), 256, 0, st>>>(R * KE, c->C, q, resp, c->wq_ws);
  LAUNCHED(c);
  StatsArgs a;
  a.B = B; a.T = T; a.D = D; a.DD = c->DD; a.diag = c->kind == SVIHMM_EMIT_NIW_DIAG; a.cat = 0; a.R = R;
  a.obs = obs; a.dtype = dtype; a.mask = mask; a.starts = starts;
  int64_t nsT = 0, nsE = 0;
  int rc;
  a.K = K; a.N = K; a.left = q; a.next = q; a.wrap = (flags & SVIHMM_WRAP) ? 1 : 0;
  if ((rc = launch_kstats(c, a, K, 0, K, &c->part_ws, &c->cap_part, &nsT, st))) return rc;
  a.K = KE; a.N = KE + 1 + D + c->DD; a.left = c->wq_ws; a.next = c->wq_ws; a.wrap = 0;
  if ((rc = launch_kstats(c, a, KE, KE, a.N, &c->part2_ws, &c->cap_part2, &nsE, st))) return rc;
  k_stats_finalize_mix<<<(unsigned)((c->slen + 255) / 256), 256, 0, st>>>(
      B, T, K, KE, D, c->DD, (int)nsT, (int)nsE, c->part_ws, c->part2_ws, q, c->seq_ws, c->prior_tran,
      (flags & SVIHMM_ADD_PRIOR) ? 1 : 0, stats_out, c->slen);
  LAUNCHED(c);
  return SVIHMM_OK;
}

// symmetric register-blocked statistics (wide64.cuh) of a dense (B, T, K) table of marginals
static int stats_sym_phase(svihmm_ctx* c, const void* obs, int dtype, const uint8_t* mask, const int64_t* starts,
                           int B, int T, const float* q, double* stats_out, unsigned flags, cudaStream_t st) {
  const int K = c->K, D = c->D;
  const int64_t R = (int64_t)B * T;
  StatsSymArgs sa;
  sa.B = B; sa.T = T; sa.K = K; sa.D = D; sa.diag = c->kind == SVIHMM_EMIT_NIW_DIAG;
  sa.NF = K + 1 + D + (sa.diag ? D : D * (D + 1) / 2);
  sa.wrap = (flags & SVIHMM_WRAP) ? 1 : 0; sa.dtype = dtype; sa.R = R;
  sa.q = q; sa.obs = obs; sa.mask = mask; sa.starts = starts;
  int64_t nsplit = std::max<int64_t>(1, std::min<int64_t>((R + 1023) / 1024, 4096));
  const int64_t rps = (((R + nsplit - 1) / nsplit) + SS_RC - 1) / SS_RC * SS_RC;
  nsplit = (R + rps - 1) / rps;
  sa.rows_per_split = rps;
  const size_t need_part = (size_t)nsplit * K * sa.NF;
  if (need_part > c->cap_part) {
    if (c->part_ws) CU(cudaFree(c->part_ws));
    c->part_ws = nullptr; c->cap_part = 0;
    CU(dalloc(&c->part_ws, need_part));
    c->cap_part = need_part;
  }
  sa.part = c->part_ws;
  static const bool no_mma = getenv("SVIHMM_STATS_FFMA") != nullptr;      // A/B switch: FFMA version
  if (no_mma) {
    dim3 grid((sa.NF + SS_TN - 1) / SS_TN, (unsigned)nsplit);
    k_stats_sym<<<grid, SS_NT, 0, st>>>(sa);
  } else {
    dim3 grid((sa.NF + SM_NT * 8 - 1) / (SM_NT * 8), (unsigned)nsplit);
    static bool attr_set = false;
    if (!attr_set) { CU(cudaFuncSetAttribute(k_stats_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_SMEM)); attr_set = true; }
    k_stats_mma<<<grid, SS_NT, SM_SMEM, st>>>(sa);
  }
  LAUNCHED(c);
  k_stats_sym_finalize<<<(unsigned)((c->slen + 255) / 256), 256, 0, st>>>(
      B, T, K, D, c->DD, sa.NF, sa.diag, (int)nsplit, c->part_ws, q, c->seq_ws, c->prior_tran,
      (flags & SVIHMM_ADD_PRIOR) ? 1 : 0, stats_out, c->slen);
  LAUNCHED(c);
  return SVIHMM_OK;
}

// ---- tensor-core statistics with TMA-fed operands (stats_tc.cuh) ------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 tmap_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
    (void)cudaGetLastError();
  }
  return fn;
}
// row-major (rows, cols) float32 matrix, box = (box_rows, cols), no swizzle, out-of-bounds rows read as zero
static bool tmap_2d_f32(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  PFN_cuTensorMapEncodeTiled_v12000 enc = tmap_encoder();
  if (!enc) return false;
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstr[1] = {cols * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool stats_tc_eligible(const svihmm_ctx* c, int T, int dtype, const void* obs, const float* q, int64_t series_rows) {
  static const bool off = getenv("SVIHMM_NO_STATS_TC") != nullptr;       // A/B switch, read once
  const int KEm = c->C > 1 ? c->KE : 0;
  const int NF = c->K + 1 + c->D + (c->kind == SVIHMM_EMIT_NIW_DIAG ? c->D : c->D * (c->D + 1) / 2);
  const int N = KEm ? (c->K + KEm + 15) / 16 * 16 : 64;
  return !off && c->K > 16 && c->K <= 64 && (c->K & 3) == 0 && (KEm & 3) == 0 && N <= 256 && ((NF + 127) / 128) * N <= 512 &&
         c->D <= 32 && (c->D & 3) == 0 && c->kind != SVIHMM_EMIT_CATEGORICAL && dtype == SVIHMM_F32 && T >= 32 &&
         ((uintptr_t)obs & 15) == 0 && ((uintptr_t)q & 15) == 0 && series_rows < (int64_t)0x7fffffff && NF <= 640 &&
         tmap_encoder() != nullptr;
}

// wq: (B*T, KE) component weights q r (mixtures) or nullptr
static int stats_tc_phase(svihmm_ctx* c, const void* obs, int64_t series_rows, const uint8_t* mask, const int64_t* starts,
                          int B, int T, const float* q, const float* wq, double* stats_out, unsigned flags, cudaStream_t st) {
  const int K = c->K, D = c->D;
  StcArgs a;
  a.B = B; a.T = T; a.K = K; a.D = D; a.diag = c->kind == SVIHMM_EMIT_NIW_DIAG;
  a.NF = K + 1 + D + (a.diag ? D : D * (D + 1) / 2);
  a.wrap = (flags & SVIHMM_WRAP) ? 1 : 0;
  a.KE = wq ? c->KE : 0;
  a.RT = wq ? 64 : 128;
  a.N = wq ? (K + a.KE + 15) / 16 * 16 : 64;
  a.ntpw = (T + a.RT - 1) / a.RT; a.nmt = (a.NF + 127) / 128;
  const int64_t nt = (int64_t)B * a.ntpw;
  if (nt >= (int64_t)0x7fffffff || (int64_t)B * T >= (int64_t)0x7fffffff) return fail(SVIHMM_EUNSUPPORTED, "minibatch too large for the tensor-core statistics");
  a.ntiles = (int)nt;
  a.q = q; a.mask = mask; a.starts = starts;
  const int grid = (int)std::min<int64_t>(148, nt);
  const int Kout = K + a.KE;
  const size_t need_part = (size_t)grid * Kout * a.NF;
  if (need_part > c->cap_part) {
    if (c->part_ws) CU(cudaFree(c->part_ws));
    c->part_ws = nullptr; c->cap_part = 0;
    CU(dalloc(&c->part_ws, need_part));
    c->cap_part = need_part;
  }
  a.part = c->part_ws;
  CUtensorMap tm_x, tm_q, tm_w;
  if (!tmap_2d_f32(&tm_x, obs, (uint64_t)series_rows, (uint64_t)D, a.RT) ||
      !tmap_2d_f32(&tm_q, q, (uint64_t)B * T, (uint64_t)K, a.RT + 1) ||
      !tmap_2d_f32(&tm_w, wq ? wq : q, (uint64_t)B * T, (uint64_t)(wq ? a.KE : K), a.RT))
    return fail(SVIHMM_ECUDA, "cuTensorMapEncodeTiled failed");
  const StcSmem L = stc_layout(K, D, a.KE, a.RT, a.N);
  const int dyn_max = c->max_smem_optin - 2048;       // the kernel also has ~1 KB of static shared memory
  if (L.total > (size_t)dyn_max) return fail(SVIHMM_EUNSUPPORTED, "tensor-core statistics: shared memory");
  static bool attr_set = false;
  if (!attr_set) { CU(cudaFuncSetAttribute(k_stats_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max)); attr_set = true; }
  k_stats_tc<<<grid, STC_NT, L.total, st>>>(tm_x, tm_q, tm_w, a);
  LAUNCHED(c);
  if (wq)
    k_stats_tc_finalize_mix<<<(unsigned)((c->slen + 255) / 256), 256, 0, st>>>(
        B, T, K, a.KE, D, c->DD, a.NF, a.diag, grid, c->part_ws, q, c->seq_ws, c->prior_tran,
        (flags & SVIHMM_ADD_PRIOR) ? 1 : 0, stats_out, c->slen);
  else
    k_stats_sym_finalize<<<(unsigned)((c->slen + 255) / 256), 256, 0, st>>>(
        B, T, K, D, c->DD, a.NF, a.diag, grid, c->part_ws, q, c->seq_ws, c->prior_tran,
        (flags & SVIHMM_ADD_PRIOR) ? 1 : 0, stats_out, c->slen);
  LAUNCHED(c);
  return SVIHMM_OK;
}

// ---- tensor-core emissions (emit_tc.cuh) ---------------------------------------------------------------
// row-major (rows, cols) float32 matrix, box = (box_rows, cols), hardware swizzle of the 16-byte chunks of a
// row (cols * 4 = 128 or 64 bytes), out-of-bounds rows read as zero
static bool tmap_2d_f32_sw(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  PFN_cuTensorMapEncodeTiled_v12000 enc = tmap_encoder();
  if (!enc) return false;
  const cuuint64_t gdim[2] = {cols, rows};
  const cuuint64_t gstr[1] = {cols * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = cols * 4 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool emit_tc_eligible(const svihmm_ctx* c, int T, int dtype, const void* obs, int64_t series_rows, unsigned flags) {
  static const bool off = getenv("SVIHMM_NO_EMIT_TC") != nullptr;        // A/B switch, read once
  const bool mix = c->C > 1;
  // D = 16 works (k_emit_tc<16>) but does not pay: the float64 kernel's work grows with D^2, this one's with D,
  // and at D = 16 they meet (BASELINE config 5: 0.89 ms either way)
  static const bool d16 = getenv("SVIHMM_EMIT_TC_D16") != nullptr;
  return !off && c->kind == SVIHMM_EMIT_NIW_FULL && (c->D == 32 || (d16 && c->D == 16)) && dtype == SVIHMM_F32 &&
         ((uintptr_t)obs & 15) == 0 && series_rows < (int64_t)0x7fffffff && T >= 32 &&
         (mix || (c->K <= 64 && (c->K & 7) == 0 && !(flags & SVIHMM_KEEP_LOCALS))) && tmap_encoder() != nullptr;
}

// K1 on the tensor cores: b / row maxima (plain emissions) or the component log-likelihoods (mixtures)
static int emit_tc_phase(svihmm_ctx* c, const void* obs, int64_t series_rows, const uint8_t* mask, const int64_t* starts,
                         int B, int T, int mask_ll, cudaStream_t st) {
  const int D = c->D, Ke = c->KE;
  const bool mix = c->C > 1;
  const int spw = 32 / D;                               // states per chunk and warpgroup
  EtcArgs a;
  a.B = B; a.T = T; a.K = Ke; a.mask_ll = mask_ll; a.zero = 0;
  a.ntpw = (T + ETC_RT - 1) / ETC_RT;
  const int64_t nt = (int64_t)B * a.ntpw;
  if (nt >= (int64_t)0x7fffffff || (int64_t)B * T >= (int64_t)0x7fffffff) return fail(SVIHMM_EUNSUPPORTED, "minibatch too large for the tensor-core emissions");
  a.ntiles = (int)nt;
  a.nchunks = ((Ke + 1) / 2 + spw - 1) / spw;
  a.H = a.nchunks * spw;
  const size_t need = (size_t)a.nchunks * etc_slot(D);
  if (need > c->cap_etc) {
    if (c->etc_blob) CU(cudaFree(c->etc_blob));
    c->etc_blob = nullptr; c->cap_etc = 0;
    CU(dalloc(&c->etc_blob, need));
    c->cap_etc = need;
  }
  a.starts = starts; a.mask = mask; a.blob = c->etc_blob; a.ck = c->ck;
  a.bout = mix ? nullptr : c->b_ws; a.mx = c->mx_ws; a.ll = mix ? c->ell_ws : nullptr;
  static const bool dbg = getenv("SVIHMM_ETC_DBG") != nullptr;           // timeline probe (scripts/emit_tc_probe.py)
  static long long* dbg_buf = nullptr;
  a.dbg = nullptr;
  if (dbg) {
    if (!dbg_buf) CU(cudaMalloc(&dbg_buf, 256 * 8 * sizeof(long long)));
    CU(cudaMemsetAsync(dbg_buf, 0, 256 * 8 * sizeof(long long), st));
    a.dbg = dbg_buf;
  }
  CUtensorMap tm_x;
  if (!tmap_2d_f32_sw(&tm_x, obs, (uint64_t)series_rows, (uint64_t)D, ETC_RT)) return fail(SVIHMM_ECUDA, "cuTensorMapEncodeTiled failed");
  const EtcSmem L = etc_layout(D, Ke);
  const int dyn_max = c->max_smem_optin - 1024;
  if (L.total > (size_t)dyn_max) return fail(SVIHMM_EUNSUPPORTED, "tensor-core emissions: shared memory");
  const int grid = (int)std::min<int64_t>(148, nt);
  static bool attr_set = false;
  if (!attr_set) {
    CU(cudaFuncSetAttribute(k_emit_tc<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max));
    CU(cudaFuncSetAttribute(k_emit_tc<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max));
    attr_set = true;
  }
  if (D == 32) {
    k_etc_prep<32><<<a.nchunks, ETC_NCOL, 0, st>>>(Ke, a.H, c->Rs, c->gk, c->etc_blob);
    LAUNCHED(c);
    k_emit_tc<32><<<grid, ETC_NT, L.total, st>>>(tm_x, a);
  } else {
    k_etc_prep<16><<<a.nchunks, ETC_NCOL, 0, st>>>(Ke, a.H, c->Rs, c->gk, c->etc_blob);
    LAUNCHED(c);
    k_emit_tc<16><<<grid, ETC_NT, L.total, st>>>(tm_x, a);
  }
  LAUNCHED(c);
  if (dbg) {
    std::vector<long long> h(256 * 8);
    CU(cudaStreamSynchronize(st));
    CU(cudaMemcpy(h.data(), dbg_buf, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    const long long t0 = h[4];
    fprintf(stderr, "k_emit_tc timeline (CTA 0, second tile; cycles since the epilogue reached chunk 0)\n"
                    "chunk  tma_issue  mma_b_ready  mma_stage_free  mma_issued | epi_arrive  epi_tfull  epi_phase1  epi_phase2\n");
    for (int ch = 0; ch < a.nchunks && ch < 256; ++ch) {
      fprintf(stderr, "%5d", ch);
      for (int i = 0; i < 8; ++i) fprintf(stderr, " %10lld%s", h[ch * 8 + i] ? h[ch * 8 + i] - t0 : 0, i == 3 ? " |" : "");
      fprintf(stderr, "\n");
    }
  }
  return SVIHMM_OK;
}

// ---- tensor-core diagonal emissions of the bf16 dense path (emit_dense.cuh) ---------------------------
static bool emit_dense_eligible(const svihmm_ctx* c, int T, int dtype, const void* obs, int64_t series_rows, unsigned flags) {
  static const bool off = getenv("SVIHMM_NO_EMIT_DENSE_TC") != nullptr;   // A/B switch, read once
  const int K = c->K, D = c->D;
  if (off || !(flags & SVIHMM_BF16_DENSE) || (flags & (SVIHMM_EXACT_XI | SVIHMM_KEEP_LOCALS)) || c->C > 1) return false;
  if (c->kind != SVIHMM_EMIT_NIW_DIAG || K <= 64 || K > 256 || (K & 3) || D < 16 || D > 64 || (D & 15)) return false;
  if (dtype != SVIHMM_F32 || ((uintptr_t)obs & 15) || series_rows >= (int64_t)0x7fffffff || T < 32 || !tmap_encoder()) return false;
  const int N = (K + 15) / 16 * 16;
  return edt_layout(D, N).total <= (size_t)c->max_smem_optin;
}
static int emit_dense_phase(svihmm_ctx* c, const void* obs, int64_t series_rows, const uint8_t* mask, const int64_t* starts,
                            int B, int T, int mask_ll, cudaStream_t st) {
  EdtArgs a;
  a.B = B; a.T = T; a.K = c->K; a.D = c->D; a.N = (c->K + 15) / 16 * 16; a.mask_ll = mask_ll;
  a.ntpw = (T + EDT_RT - 1) / EDT_RT;
  const int64_t nt = (int64_t)B * a.ntpw;
  if (nt >= (int64_t)0x7fffffff || (int64_t)B * T >= (int64_t)0x7fffffff) return fail(SVIHMM_EUNSUPPORTED, "minibatch too large for the tensor-core emissions");
  a.ntiles = (int)nt;
  a.starts = starts; a.mask = mask; a.par2 = c->par2; a.ckp = c->ckp; a.bout = c->b_ws; a.mx = c->mx_ws;
  CUtensorMap tm_x;
  if (!tmap_2d_f32(&tm_x, obs, (uint64_t)series_rows, (uint64_t)c->D, EDT_RT)) return fail(SVIHMM_ECUDA, "cuTensorMapEncodeTiled failed");
  const EdtSmem L = edt_layout(c->D, a.N);
  static bool attr_set = false;
  if (!attr_set) { CU(cudaFuncSetAttribute(k_emit_diag_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, c->max_smem_optin)); attr_set = true; }
  k_emit_diag_tc<<<(int)std::min<int64_t>(148, nt), EDT_NT, L.total, st>>>(tm_x, a);
  LAUNCHED(c);
  return SVIHMM_OK;
}

// generic statistics contraction (stats.cuh) of a dense (B, T, K) table of marginals
static int stats_generic_phase(svihmm_ctx* c, const void* obs, int dtype, const uint8_t* mask,
                               const int64_t* starts, int B, int T, const float* q, double* stats_out,
                               unsigned flags, bool xi, cudaStream_t st, const uint16_t* q16 = nullptr) {
  const int K = c->K, D = c->D;
  const int64_t R = (int64_t)B * T;
  // K4: statistics
  StatsArgs a;
  a.B = B; a.T = T; a.K = K; a.D = D; a.DD = c->DD; a.N = c->nfeat;
  a.diag = c->kind == SVIHMM_EMIT_NIW_DIAG; a.cat = c->kind == SVIHMM_EMIT_CATEGORICAL; a.R = R;
  a.obs = obs; a.dtype = dtype; a.mask = mask; a.starts = starts;
  const int TM = K <= 16 ? 16 : (K <= 32 ? 32 : 64);
  const int tiles_m = (K + TM - 1) / TM;
  const int tiles_n = (c->nfeat + ST_TN - 1) / ST_TN;
  const int64_t chunks = (R + ST_RC - 1) / ST_RC;
  int64_t nsplit = (4 * 148 + (int64_t)tiles_m * tiles_n - 1) / ((int64_t)tiles_m * tiles_n);
  if (nsplit > chunks) nsplit = chunks;
  if (nsplit < 1) nsplit = 1;
  int64_t rps = ((chunks + nsplit - 1) / nsplit) * ST_RC;
  nsplit = (R + rps - 1) / rps;
  a.rows_per_split = rps;
  const size_t need_part = (size_t)nsplit * K * c->nfeat;
  if (need_part > c->cap_part) {
    if (c->part_ws) CU(cudaFree(c->part_ws));
    c->part_ws = nullptr; c->cap_part = 0;
    CU(dalloc(&c->part_ws, need_part));
    c->cap_part = need_part;
  }
  a.part = c->part_ws;
  const size_t xsm = (size_t)ST_RC * D * sizeof(float);
  auto launch_stats = [&](const float* left, const float* next, int n_lo, int n_hi, int wrap) {
    a.left = left; a.next = next; a.n_lo = n_lo; a.n_hi = n_hi; a.wrap = wrap;
    dim3 grid((n_hi - n_lo + ST_TN - 1) / ST_TN, tiles_m, (unsigned)nsplit);
    if (TM == 16) k_stats<16><<<grid, 256, xsm, st>>>(a);
    else if (TM == 32) k_stats<32><<<grid, 256, xsm, st>>>(a);
    else k_stats<64><<<grid, 256, xsm, st>>>(a);
    c->launches++;
  };
  int nsplit_fin = (int)nsplit;
  if (q16) {
    // dense path: the K x K transition block on the tensor cores from the bf16 marginals in tile layout
    // (k_tran_stats_dense adds into split 0 of the zeroed partials); the emission columns likewise when
    // the feature rows fit one operand slot (k_emit_stats_dense), else by k_stats
    const int KPd = (K + 63) / 64 * 64, tiles = (B + DN_M - 1) / DN_M;
    const int NB = c->nfeat - K;
    static const bool ffma_emit_stats = getenv("SVIHMM_DENSE_FFMA_EMIT_STATS") != nullptr;   // A/B switch, read once
    const bool tc_emit = !a.cat && NB <= ESD_NB && !ffma_emit_stats;
    const int NP = (flags & SVIHMM_WRAP) ? T : T - 1;
    const int TS = std::max(1, std::min(NP, 148 / tiles));
    const int TSe = std::max(1, std::min(T, 148 / tiles));
    // every CTA of the tensor-core kernels owns one split of the partials (no float atomics: the sums are
    // run-to-run deterministic); unused (split, column block) combinations stay zero
    const int nsd = std::max(tiles * TS, tc_emit ? tiles * TSe : (int)nsplit);
    const size_t need_d = (size_t)nsd * K * c->nfeat;
    if (need_d > c->cap_part) {
      if (c->part_ws) CU(cudaFree(c->part_ws));
      c->part_ws = nullptr; c->cap_part = 0;
      CU(dalloc(&c->part_ws, need_d));
      c->cap_part = need_d;
      a.part = c->part_ws;
    }
    CU(cudaMemsetAsync(c->part_ws, 0, need_d * sizeof(float), st));
    nsplit_fin = nsd;
    if (tc_emit) {
      const size_t needf = (size_t)tiles * T * NB * DN_M;
      if (needf > c->cap_dnf) {
        if (c->dn_fhi) CU(cudaFree(c->dn_fhi));
        if (c->dn_flo) CU(cudaFree(c->dn_flo));
        c->dn_fhi = c->dn_flo = nullptr; c->cap_dnf = 0;
        CU(dalloc(&c->dn_fhi, needf)); CU(dalloc(&c->dn_flo, needf));
        c->cap_dnf = needf;
      }
      k_dense_tile_feat<<<dim3(T, tiles), DN_M, 0, st>>>(B, T, D, NB, a.diag, obs, dtype, mask, starts,
                                                        reinterpret_cast<__nv_bfloat16*>(c->dn_fhi),
                                                        reinterpret_cast<__nv_bfloat16*>(c->dn_flo));
      LAUNCHED(c);
      const size_t smem_e = (size_t)TSD_SLOT + 2 * (size_t)ESD_NB * 256 + 1024;
      CU(cudaFuncSetAttribute(k_emit_stats_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_e));
      k_emit_stats_dense<<<dim3(tiles, TSe), 256, smem_e, st>>>(T, K, KPd, NB, TSe, reinterpret_cast<const __nv_bfloat16*>(q16),
                                                               reinterpret_cast<const __nv_bfloat16*>(c->dn_fhi),
                                                               reinterpret_cast<const __nv_bfloat16*>(c->dn_flo),
                                                               c->part_ws, c->nfeat, K);
      LAUNCHED(c);
    } else {
      launch_stats(q, q, K, c->nfeat, 0);
    }
    const size_t smem = 3 * (size_t)TSD_SLOT + 1024;
    CU(cudaFuncSetAttribute(k_tran_stats_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (NP > 0) {
      k_tran_stats_dense<<<dim3(tiles, TS), 256, smem, st>>>(T, K, KPd, NP, TS, reinterpret_cast<const __nv_bfloat16*>(q16),
                                                            c->part_ws, c->nfeat);
      LAUNCHED(c);
    }
  } else if (xi) {
    launch_stats(c->alpha_ws, c->r_ws, 0, K, 0);
    launch_stats(q, q, K, c->nfeat, 0);
  } else {
    launch_stats(q, q, 0, c->nfeat, (flags & SVIHMM_WRAP) ? 1 : 0);
  }
  CU(cudaGetLastError());
  k_stats_finalize<<<(unsigned)((c->slen + 255) / 256), 256, 0, st>>>(
      B, T, K, D, c->DD, c->nfeat, nsplit_fin, c->part_ws, q, c->seq_ws, c->prior_tran,
      (flags & SVIHMM_ADD_PRIOR) ? 1 : 0, c->Pt, xi ? 1 : 0, stats_out, c->slen);
  LAUNCHED(c);
  return SVIHMM_OK;
}

// Batched tensor-core E-step (batch16.cuh): the diagonal model with K <= 16, D <= 16; sixteen windows
// per chain warp, so it is the path for minibatches (the pipelined one-CTA-per-window kernel remains for
// callers that set the threshold above their B through svihmm_set_tuning).
static bool b16_eligible(const svihmm_ctx* c, int B, unsigned flags) {
  return c->b16_min_B > 0 && B >= c->b16_min_B && c->K <= 16 && c->C == 1 && c->kind == SVIHMM_EMIT_NIW_DIAG &&
         c->D <= 16 && !(flags & (SVIHMM_EXACT_XI | SVIHMM_KEEP_LOCALS));
}

static int estep_b16(svihmm_ctx* c, const void* obs, int dtype, const uint8_t* mask, const int64_t* starts,
                     int B, int T, float* var_x_out, double* stats_out, unsigned flags, cudaStream_t st) {
  const int K = c->K, D = c->D;
  const size_t rows = (size_t)B * T;
  if (rows > c->cap_b16) {
    void* olds[] = {c->b16_b, c->b16_a, c->b16_c, c->b16_E, c->b16_mx};
    for (void* p : olds) if (p) CU(cudaFree(p));
    c->b16_b = c->b16_a = c->b16_c = nullptr; c->b16_E = nullptr; c->b16_mx = nullptr; c->cap_b16 = 0;
    CU(dalloc(&c->b16_b, rows * B16_KS)); CU(dalloc(&c->b16_a, rows * B16_KS)); CU(dalloc(&c->b16_c, rows * B16_KS));
    CU(dalloc(&c->b16_E, rows)); CU(dalloc(&c->b16_mx, rows));
    c->cap_b16 = rows;
  }
  if ((size_t)B > c->cap_B) {
    if (c->seq_ws) CU(cudaFree(c->seq_ws));
    c->seq_ws = nullptr; c->cap_B = 0;
    CU(dalloc(&c->seq_ws, 2 * (size_t)B));
    c->cap_B = B;
  }
  B16Args a;
  a.B = B; a.T = T; a.K = K; a.D = D;
  a.wrap = (flags & SVIHMM_WRAP) ? 1 : 0; a.add_prior = (flags & SVIHMM_ADD_PRIOR) ? 1 : 0;
  a.mask_ll = (flags & SVIHMM_MASK_LL) ? 1 : 0;
  a.obs = obs; a.dtype = dtype; a.mask = mask; a.starts = starts;
  a.P = c->Pt; a.PT = c->PtT; a.pi0 = c->pi0; a.par2 = c->par2; a.ckp = c->ckp; a.prior_tran = c->prior_tran;
  a.bt = c->b16_b; a.at = c->b16_a; a.ct = c->b16_c; a.Et = c->b16_E; a.mx = c->b16_mx;
  a.var_x_out = var_x_out; a.stats_out = stats_out; a.seq = c->seq_ws;
  a.o_n = (size_t)K * K; a.o_sx = a.o_n + K; a.o_sxx = a.o_sx + (size_t)K * D;
  a.o_q0 = a.o_sxx + (size_t)K * c->DD; a.o_tail = a.o_q0 + K; a.slen = c->slen;
  static const bool dbg = getenv("SVIHMM_B16_DBG") != nullptr;
#define B16_DBG(what) do { if (dbg) { cudaError_t e_ = cudaStreamSynchronize(st); \
    fprintf(stderr, "[b16 dbg] %s done: %s (B=%d T=%d K=%d D=%d)\n", what, cudaGetErrorString(e_), B, T, K, D); fflush(stderr); } } while (0)
  {
    PhaseTimer pt(c, PH_EMIT, st);
    const unsigned grid = (unsigned)std::min<size_t>((rows + 255) / 256, 148 * 16);
    const int KPe = K <= 4 ? 4 : (K <= 8 ? 8 : 16);
    const size_t smem = (size_t)(2 * KPe * D + KPe) * sizeof(double);
    if (KPe == 4) k_b16_emit<4><<<grid, 256, smem, st>>>(a);
    else if (KPe == 8) k_b16_emit<8><<<grid, 256, smem, st>>>(a);
    else k_b16_emit<16><<<grid, 256, smem, st>>>(a);
    LAUNCHED(c);
  }
  B16_DBG("emit");
  {
    PhaseTimer pt(c, PH_FORWARD, st);
    const int ngroups = (B + 15) / 16, nwarps = 2 * ngroups;
    // few chain warps: one per CTA so that each has an SM (scheduler, tensor pipe) to itself
    const int wpb = nwarps <= 148 ? 1 : 4;
    k_b16_chain<<<(nwarps + wpb - 1) / wpb, wpb * 32, 0, st>>>(a, ngroups);
    LAUNCHED(c);
  }
  B16_DBG("chain");
  {
    PhaseTimer pt(c, PH_STATS, st);
    const int64_t nunits = (int64_t)B * ((T + 7) / 8);
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>(2 * 148, (nunits + 7) / 8));
    if (D <= 8) k_b16_post<3><<<grid, 256, 0, st>>>(a);
    else k_b16_post<5><<<grid, 256, 0, st>>>(a);
    LAUNCHED(c);
  }
  B16_DBG("post");
#undef B16_DBG
  c->last_B = B; c->last_T = T; c->last_fused = 1;
  return SVIHMM_OK;
}

static int estep_impl(svihmm_ctx* c, const void* obs, int dtype, const uint8_t* mask,
                      const int64_t* starts, int B, int T, float* var_x_out, double* stats_out,
                      unsigned flags, cudaStream_t st, int trim = 0) {
  const int K = c->K, D = c->D;
  const bool xi = flags & SVIHMM_EXACT_XI;
  // rows behind `obs`: the resident series, or the B*T gathered rows of a staging buffer
  const int64_t series_rows = obs == c->obs ? c->T_full : (int64_t)B * T;
  if (trim < 0 || 2 * trim >= T) return fail(SVIHMM_EINVAL, "trim = %d leaves no inner rows of T = %d", trim, T);
  if (trim > 0 && xi) return fail(SVIHMM_EUNSUPPORTED, "SVIHMM_EXACT_XI with buffered windows");
  size_t fsmem = 0;
  if (trim > 0) {}    // buffered windows: per-phase kernels (the statistics see the compacted inner rows)
  else if (b16_eligible(c, B, flags))
    return estep_b16(c, obs, dtype, mask, starts, B, T, var_x_out, stats_out, flags, st);
  else if (pipe_eligible(c, T, flags, &fsmem))
    return estep_fused(c, obs, dtype, mask, starts, B, T, var_x_out, stats_out, flags, fsmem, st, true);
  else if (fused_eligible(c, T, flags, &fsmem))
    return estep_fused(c, obs, dtype, mask, starts, B, T, var_x_out, stats_out, flags, fsmem, st, false);
  int rc = ensure_ws(c, B, T, xi);
  if (rc) return rc;
  const int64_t R = (int64_t)B * T;
  const int mask_ll = (flags & SVIHMM_MASK_LL) ? 1 : 0;
  // K1: expected log-likelihoods (fp64) -> scaled likelihoods b (fp32) + row maxima.  Mixtures:
  // the K*C component log-likelihoods first, then the per-state logsumexp + responsibilities.
  const bool mix = c->C > 1;
  if (mix && xi) return fail(SVIHMM_EUNSUPPORTED, "SVIHMM_EXACT_XI with mixture emissions");
  if (mix) {
    const size_t rows = (size_t)B * T;
    if (rows > c->cap_rows_mix) {
      void* olds[] = {c->ell_ws, c->resp_ws, c->wq_ws};
      for (void* p : olds) if (p) CU(cudaFree(p));
      c->ell_ws = nullptr; c->resp_ws = nullptr; c->wq_ws = nullptr; c->cap_rows_mix = 0;
      CU(dalloc(&c->ell_ws, rows * c->KE)); CU(dalloc(&c->resp_ws, rows * c->KE)); CU(dalloc(&c->wq_ws, rows * c->KE));
      c->cap_rows_mix = rows;
    }
  }
  { PhaseTimer pt(c, PH_EMIT, st);
  const int Ke = c->KE;                                   // emission components (= K without mixtures)
  double* ll_out = mix ? c->ell_ws : c->ll_ws;
  float* b_out = mix ? nullptr : c->b_ws;
  if (emit_dense_eligible(c, T, dtype, obs, series_rows, flags)) {
    // bf16 dense path: diagonal emissions as one bf16 hi + lo contraction on tcgen05 (emit_dense.cuh)
    if ((rc = emit_dense_phase(c, obs, series_rows, mask, starts, B, T, mask_ll, st))) return rc;
  } else if (emit_tc_eligible(c, T, dtype, obs, series_rows, flags)) {
    // exact sliced product on tcgen05, float64 square-and-sum epilogue (emit_tc.cuh)
    if ((rc = emit_tc_phase(c, obs, series_rows, mask, starts, B, T, mask_ll, st))) return rc;
  } else if (c->kind == SVIHMM_EMIT_NIW_FULL && (D == 8 || D == 16 || D == 32)) {
    // register-blocked float64 kernel with the row maximum and b = exp(ll - max) fused in
    // (a persistent-grid variant that kept the float64 log-likelihoods in per-CTA scratch slots, so that
    // the 268 MB table of config 3 never reaches HBM, was measured SLOWER: 2.69 ms against 1.87 ms; the
    // kernel is bound by the FP64 pipe, not by that write, and loses its CTA-level latency hiding)
    const unsigned grid = (unsigned)((R + 2 * ERB_NT - 1) / (2 * ERB_NT));
#define ERB_LAUNCH(DV) do { \
      const size_t smem = (size_t)2 * ERB_KC * (erb_len(DV) + DV + (DV & 1)) * sizeof(double); \
      if (smem > 48 * 1024) CU(cudaFuncSetAttribute(k_emit_full_rb<DV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      k_emit_full_rb<DV><<<grid, ERB_NT, smem, st>>>(R, T, Ke, obs, dtype, mask, starts, mask_ll, c->Rs, c->gk, \
                                                     c->ck, ll_out, b_out, c->mx_ws); } while (0)
    if (D == 8) ERB_LAUNCH(8); else if (D == 16) ERB_LAUNCH(16); else ERB_LAUNCH(32);
#undef ERB_LAUNCH
    LAUNCHED(c);
  } else if (c->kind == SVIHMM_EMIT_CATEGORICAL) {
    k_emit_cat<<<(unsigned)((R * K + 255) / 256), 256, 0, st>>>(B, T, K, D, obs, dtype, mask, starts, mask_ll,
                                                             c->Rs, c->ll_ws);
    LAUNCHED(c);
    k_ll_to_b<<<(unsigned)((R * 32 + 255) / 256), 256, 0, st>>>(R, K, c->ll_ws, c->b_ws, c->mx_ws);
    LAUNCHED(c);
  } else {
  bool diag_fused_b = false;
  if (c->kind == SVIHMM_EMIT_NIW_FULL) {
    const size_t smem = ((size_t)EMIT_ROWS * D + (size_t)D * (D + 1) / 2 + D) * sizeof(double);
    if (smem > 200 * 1024) return fail(SVIHMM_EUNSUPPORTED, "D = %d too large for the emission kernel", D);
    if (smem > 48 * 1024) CU(cudaFuncSetAttribute(k_emit_full, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_emit_full<<<(unsigned)((R + EMIT_ROWS - 1) / EMIT_ROWS), EMIT_ROWS, smem, st>>>(
        B, T, Ke, D, obs, dtype, mask, starts, mask_ll, c->Rs, c->gk, c->ck, ll_out);
  } else {
    size_t smem = 2 * (size_t)Ke * D * sizeof(double);
    if (Ke >= 32 && (D == 64 || D == 32 || D == 16)) {
      // register-blocked float64 kernel (row in registers), row maximum and b = exp(ll - max) fused in
      const unsigned grid = (unsigned)((R + EDR_NT - 1) / EDR_NT);
#define EDR_LAUNCH(DV) do { \
      const size_t sm2 = (size_t)2 * EDR_KC * DV * sizeof(double2); \
      if (sm2 > 48 * 1024) CU(cudaFuncSetAttribute(k_emit_diag_rb<DV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2)); \
      k_emit_diag_rb<DV><<<grid, EDR_NT, sm2, st>>>(R, T, Ke, obs, dtype, mask, starts, mask_ll, c->par2, c->ckp, \
                                                     ll_out, b_out, c->mx_ws); } while (0)
      if (D == 64) EDR_LAUNCH(64); else if (D == 32) EDR_LAUNCH(32); else EDR_LAUNCH(16);
#undef EDR_LAUNCH
      diag_fused_b = !mix;
    } else if (smem <= 200 * 1024) {
      if (smem > 48 * 1024) CU(cudaFuncSetAttribute(k_emit_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_emit_diag<<<(unsigned)((R * Ke + 255) / 256), 256, smem, st>>>(
          B, T, Ke, D, obs, dtype, mask, starts, mask_ll, c->Rs, c->gk, c->ck, ll_out, 1);
    } else {
      // parameters of all states do not fit: tiles of 32 states (config 4: K*D = 16384)
      smem = 2 * (size_t)32 * D * sizeof(double);
      if (smem > 200 * 1024) return fail(SVIHMM_EUNSUPPORTED, "D = %d too large for the diagonal emission kernel", D);
      if (smem > 48 * 1024) CU(cudaFuncSetAttribute(k_emit_diag_tiled, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const unsigned gx = (unsigned)std::min<int64_t>((R + 8 * ED_R - 1) / (8 * ED_R), 148 * 8);
      k_emit_diag_tiled<<<dim3(gx, (Ke + 31) / 32), 256, smem, st>>>(
          B, T, Ke, D, obs, dtype, mask, starts, mask_ll, c->Rs, c->gk, c->ck, ll_out);
    }
  }
  LAUNCHED(c);
  if (!mix && !diag_fused_b) {
    k_ll_to_b<<<(unsigned)((R * 32 + 255) / 256), 256, 0, st>>>(R, K, c->ll_ws, c->b_ws, c->mx_ws);
    LAUNCHED(c);
  }
  }
  if (mix) {
    k_mix_combine<<<(unsigned)((R * K + 255) / 256), 256, 0, st>>>(B, T, K, c->C, D, obs, dtype, mask, starts, mask_ll,
                                                                c->lw, c->ell_ws, c->ll_ws, c->resp_ws);
    LAUNCHED(c);
    k_ll_to_b<<<(unsigned)((R * 32 + 255) / 256), 256, 0, st>>>(R, K, c->ll_ws, c->b_ws, c->mx_ws);
    LAUNCHED(c);
  }
  }
  // K2/K3: forward, backward + marginals
  float* q = var_x_out ? var_x_out : c->q_ws;
  float* r = xi ? c->r_ws : nullptr;
  float* cs = (float*)(c->mx_ws + (size_t)B * T);
  // the statistics phase works on (qs, starts_s, Ts): the marginals themselves, or their inner rows
  const float* qs = q; const int64_t* starts_s = starts; int Ts = T;
  auto trim_for_stats = [&]() -> int {
    if (trim == 0) return SVIHMM_OK;
    int rc_ = trim_table(c, B, T, trim, K, q, &c->qin_ws, &c->cap_qin, st);
    if (rc_) return rc_;
    if ((size_t)B > c->cap_startsin) {
      if (c->starts_in) CU(cudaFree(c->starts_in));
      c->starts_in = nullptr; c->cap_startsin = 0;
      CU(dalloc(&c->starts_in, (size_t)B));
      c->cap_startsin = B;
    }
    k_shift_starts<<<(B + 255) / 256, 256, 0, st>>>(B, trim, starts, c->starts_in);
    LAUNCHED(c);
    qs = c->qin_ws; starts_s = c->starts_in; Ts = T - 2 * trim;
    return SVIHMM_OK;
  };
  if (K > 16 && K <= 64 && !xi && !(flags & SVIHMM_KEEP_LOCALS) && D <= 64 && c->kind != SVIHMM_EMIT_CATEGORICAL) {
    // warp-per-chain recursions, marginals, symmetric register-blocked statistics (wide64.cuh)
    if (!c->r_ws) CU(dalloc(&c->r_ws, c->cap_rows * K));
    { PhaseTimer pt(c, PH_FORWARD, st);
      k_chain_wide<<<(2 * B + 3) / 4, 128, 0, st>>>(B, T, K, c->Pt, c->PtT, c->pi0, c->b_ws, c->alpha_ws, c->r_ws, c->e_ws);
      LAUNCHED(c); }
    { PhaseTimer pt(c, PH_BACKWARD, st);
      k_marginals_wide<<<148 * 8, 256, 0, st>>>(R, K, c->alpha_ws, c->r_ws, c->e_ws, q, c->lt_ws);
      LAUNCHED(c);
      k_seq_logz_lt<<<(B * 32 + 255) / 256, 256, 0, st>>>(B, T, c->lt_ws, c->mx_ws, c->seq_ws);
      LAUNCHED(c); }
    PhaseTimer pt_stats(c, PH_STATS, st);
    if ((rc = trim_for_stats())) return rc;
    if (mix) {
      c->last_B = B; c->last_T = T; c->last_fused = 1;
      if (trim == 0 && stats_tc_eligible(c, Ts, dtype, obs, qs, series_rows)) {
        // component weights q r, then ONE tensor-core contraction for transitions and component statistics
        k_mix_weights<<<(unsigned)((R * c->KE + 255) / 256), 256, 0, st>>>(R * c->KE, c->C, qs, c->resp_ws, c->wq_ws);
        LAUNCHED(c);
        return stats_tc_phase(c, obs, series_rows, mask, starts_s, B, Ts, qs, c->wq_ws, stats_out, flags, st);
      }
      return stats_mix(c, obs, dtype, mask, starts_s, B, Ts, qs, T, trim, stats_out, flags, st);
    }
    c->last_B = B; c->last_T = T; c->last_fused = 1;     // no lliks/alpha/cs tables in the classic form
    if (stats_tc_eligible(c, Ts, dtype, obs, qs, series_rows))
      return stats_tc_phase(c, obs, series_rows, mask, starts_s, B, Ts, qs, nullptr, stats_out, flags, st);
    return stats_sym_phase(c, obs, dtype, mask, starts_s, B, Ts, qs, stats_out, flags, st);
  }
  if ((flags & SVIHMM_BF16_DENSE) && K > 64 && K <= 256 && (K & 3) == 0 && !xi && !(flags & SVIHMM_KEEP_LOCALS)) {
    // dense K x K step on tcgen05 (bf16 messages, float32 accumulators in tensor memory): dense.cuh
    const int KPd = (K + 63) / 64 * 64;
    const size_t smem = (size_t)DN_M * KPd * 2 + (size_t)KPd * KPd * 2 + 1024;
    const int tiles = (B + DN_M - 1) / DN_M;
    static const bool ffma_tran_stats = getenv("SVIHMM_DENSE_FFMA_STATS") != nullptr;         // A/B switch, read once
    const bool tc_tran = trim == 0 && !mix && !ffma_tran_stats;   // transition statistic on tcgen05
    const size_t need = (size_t)tiles * DN_M * T * K;
    if (need > c->cap_dn) {
      void* olds[] = {c->dn_b, c->dn_a, c->dn_r, c->dn_e, c->dn_q16};
      for (void* p : olds) if (p) CU(cudaFree(p));
      c->dn_b = c->dn_a = c->dn_r = nullptr; c->dn_e = nullptr; c->dn_q16 = nullptr; c->cap_dn = 0;
      CU(dalloc(&c->dn_q16, need));
      CU(dalloc(&c->dn_b, need)); CU(dalloc(&c->dn_a, need)); CU(dalloc(&c->dn_r, need));
      CU(dalloc(&c->dn_e, (size_t)tiles * DN_M * T));
      c->cap_dn = need;
    }
    CU(cudaFuncSetAttribute(k_chain_dense, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { PhaseTimer pt(c, PH_FORWARD, st);
      k_dense_tile_b<<<dim3(T, tiles), 256, 0, st>>>(B, T, K, c->b_ws, c->dn_b);
      LAUNCHED(c);
      static const bool no_cl = getenv("SVIHMM_NO_DENSE_CLUSTER") != nullptr;    // A/B switch, read once
      if (KPd == DNC_KP && !no_cl) {
        // N split over a cluster of 4 CTAs, carried vector exchanged through distributed shared memory
        const size_t smc = 2 * (size_t)DN_M * DNC_KP * 2 + (size_t)DNC_NS * DNC_KP * 2 + 2 * 16 * DN_M * sizeof(float) + 1024;
        static bool cl_attr = false;
        if (!cl_attr) { CU(cudaFuncSetAttribute(k_chain_dense_cl, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smc)); cl_attr = true; }
        k_chain_dense_cl<<<dim3(tiles * DNC_R, 2), DN_M * 4, smc, st>>>(B, T, K, c->PtT, c->Pt, c->pi0, c->dn_b,
                                                                   c->dn_a, c->dn_r, c->dn_e);
      } else {
        k_chain_dense<<<dim3(tiles, 2), DN_M * DN_NS, smem, st>>>(B, T, K, KPd, c->PtT, c->Pt, c->pi0, c->dn_b,
                                                         c->dn_a, c->dn_r, c->dn_e);
      }
      LAUNCHED(c); }
    { PhaseTimer pt(c, PH_BACKWARD, st);
      k_marginals_tiled<<<dim3(T, tiles), DN_M, 0, st>>>(B, T, K, c->dn_a, c->dn_r, c->dn_e, q, c->lt_ws,
                                                         tc_tran ? reinterpret_cast<__nv_bfloat16*>(c->dn_q16) : nullptr);
      LAUNCHED(c);
      k_seq_logz_lt<<<(B * 32 + 255) / 256, 256, 0, st>>>(B, T, c->lt_ws, c->mx_ws, c->seq_ws);
      LAUNCHED(c); }
    PhaseTimer pt_stats(c, PH_STATS, st);
    if ((rc = trim_for_stats())) return rc;
    c->last_B = B; c->last_T = T; c->last_fused = 1;
    if (mix) return stats_mix(c, obs, dtype, mask, starts_s, B, Ts, qs, T, trim, stats_out, flags, st);
    return stats_generic_phase(c, obs, dtype, mask, starts_s, B, Ts, qs, stats_out, flags, false, st,
                               tc_tran ? c->dn_q16 : nullptr);
  }
  // KEEP_LOCALS: also keep the normalised backward messages + their scale factors (self.lbeta)
  float* beta_out = nullptr; float* sb_out = nullptr;
  c->last_beta = 0;
  if (flags & SVIHMM_KEEP_LOCALS) {
    if ((size_t)R > c->cap_beta) {
      if (c->beta_ws) CU(cudaFree(c->beta_ws));
      if (c->sb_ws) CU(cudaFree(c->sb_ws));
      c->beta_ws = c->sb_ws = nullptr; c->cap_beta = 0;
      CU(dalloc(&c->beta_ws, (size_t)R * K)); CU(dalloc(&c->sb_ws, (size_t)R));
      c->cap_beta = (size_t)R;
    }
    beta_out = c->beta_ws; sb_out = c->sb_ws; c->last_beta = 1;
  }
  static const bool no_scan = getenv("SVIHMM_NO_SCAN") != nullptr;         // A/B switch, read once
  if (K <= 16 && !xi && !no_scan && c->scan_min_T > 0 && T >= c->scan_min_T && (int64_t)B * ((T + 255) / 256) >= 64) {
    // few LONG chains: block-parallel scan (scan16.cuh) instead of T sequential steps on one warp
    ScanArgs sa;
    sa.B = B; sa.T = T; sa.K = K; sa.Lc = T >= (1 << 19) ? 512 : 256; sa.C = (T + sa.Lc - 1) / sa.Lc;
    sa.P = c->Pt; sa.PT = c->PtT; sa.pi0 = c->pi0; sa.b = c->b_ws; sa.alpha = c->alpha_ws; sa.cs = cs;
    sa.q = q; sa.beta = beta_out; sa.sb = sb_out;
    const size_t nch = (size_t)B * sa.C;
    if (nch > c->cap_scan) {
      if (c->scan_ops) CU(cudaFree(c->scan_ops));
      if (c->scan_bound) CU(cudaFree(c->scan_bound));
      c->scan_ops = c->scan_bound = nullptr; c->cap_scan = 0;
      CU(dalloc(&c->scan_ops, nch * 2 * SC_OP)); CU(dalloc(&c->scan_bound, nch * 2 * 16));
      c->cap_scan = nch;
    }
    sa.ops = c->scan_ops; sa.bound = c->scan_bound;
    const int ngroups = (int)((nch + 15) / 16);
    { PhaseTimer pt(c, PH_FORWARD, st);
      k_scan_ops<<<(unsigned)((2 * nch + 3) / 4), 128, 0, st>>>(sa); LAUNCHED(c);
      k_scan_combine<<<B, 64, 0, st>>>(sa); LAUNCHED(c);
      k_scan_pass<<<(ngroups + 3) / 4, 128, 0, st>>>(sa, ngroups, 1); LAUNCHED(c); }
    { PhaseTimer pt(c, PH_BACKWARD, st);
      k_scan_pass<<<(ngroups + 3) / 4, 128, 0, st>>>(sa, ngroups, 0); LAUNCHED(c); }
  } else if (K <= 32) {
    switch (c->KP) {
      case 2: launch_fb<2>(c, B, T, q, r, st, beta_out, sb_out); break;
      case 4: launch_fb<4>(c, B, T, q, r, st, beta_out, sb_out); break;
      case 8: launch_fb<8>(c, B, T, q, r, st, beta_out, sb_out); break;
      case 16: launch_fb<16>(c, B, T, q, r, st, beta_out, sb_out); break;
      default: launch_fb<32>(c, B, T, q, r, st, beta_out, sb_out); break;
    }
  } else {
    const int KT = (K + 31) / 32 * 32;
    size_t smem = ((size_t)K * K + 2 * KT + 64) * sizeof(float);
    const int p_smem = smem <= (size_t)c->max_smem_optin;
    if (!p_smem) smem = (2 * (size_t)KT + 64) * sizeof(float);
    if (smem > 48 * 1024) {
      CU(cudaFuncSetAttribute(k_forward_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CU(cudaFuncSetAttribute(k_backward_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    { PhaseTimer pt(c, PH_FORWARD, st);
      k_forward_wide<<<B, KT, smem, st>>>(B, T, K, c->Pt, c->pi0, c->b_ws, c->alpha_ws, cs, p_smem);
      c->launches++; }
    { PhaseTimer pt(c, PH_BACKWARD, st);
      k_backward_wide<<<B, KT, smem, st>>>(B, T, K, c->PtT, c->b_ws, c->alpha_ws, q, r, p_smem, beta_out, sb_out);
      c->launches++; }
  }
  CU(cudaGetLastError());
  PhaseTimer pt_stats(c, PH_STATS, st);
  if (T >= 8192) {                    // long chains: split the rows of a sequence over CTAs
    CU(cudaMemsetAsync(c->seq_ws, 0, sizeof(double) * 2 * B, st));
    k_seq_logz_split<<<dim3(B, (unsigned)std::min(256, T / 4096)), 256, 0, st>>>(B, T, cs, c->mx_ws, c->seq_ws);
  } else {
    k_seq_logz<<<(B * 32 + 255) / 256, 256, 0, st>>>(B, T, cs, c->mx_ws, c->seq_ws);
  }
  LAUNCHED(c);
  if ((rc = trim_for_stats())) return rc;
  if (mix) {
    c->last_B = B; c->last_T = T; c->last_fused = 0;
    return stats_mix(c, obs, dtype, mask, starts_s, B, Ts, qs, T, trim, stats_out, flags, st);
  }
  c->last_B = B; c->last_T = T; c->last_fused = 0;
  return stats_generic_phase(c, obs, dtype, mask, starts_s, B, Ts, qs, stats_out, flags, xi, st);
}

static int check_estep_args(svihmm_ctx* c, const void* starts, int B, int T, const void* stats, unsigned flags) {
  if (!c || !starts || !stats) return fail(SVIHMM_EINVAL, "NULL argument");
  if (B < 1 || T < 1) return fail(SVIHMM_EINVAL, "B (%d) and T (%d) must be >= 1", B, T);
  if (!c->have_globals) return fail(SVIHMM_ESTATE, "svihmm_set_globals has not been called");
  if ((flags & SVIHMM_ADD_PRIOR) && !c->have_prior) return fail(SVIHMM_ESTATE, "SVIHMM_ADD_PRIOR needs svihmm_set_prior");
  return SVIHMM_OK;
}

extern "C" int svihmm_estep(svihmm_ctx* c, const int64_t* starts, int B, int T, float* var_x_out,
                            double* stats_out, unsigned flags, void* stream) {
  int rc = check_estep_args(c, starts, B, T, stats_out, flags);
  if (rc) return rc;
  if (!c->obs) return fail(SVIHMM_ESTATE, "svihmm_set_series has not been called");
  if (T > c->T_full) return fail(SVIHMM_EINVAL, "T (%d) exceeds the series length (%lld)", T, (long long)c->T_full);
  CU(cudaSetDevice(c->device));
  return estep_impl(c, c->obs, c->obs_dtype, c->mask, starts, B, T, var_x_out, stats_out, flags,
                    (cudaStream_t)stream);
}

/* buffered meta-observations + explicit initial-state parameter (adaptive window machinery) */
extern "C" int svihmm_estep_buffered(svihmm_ctx* c, const int64_t* starts, int B, int T, int trim,
                                     float* var_x_out, double* stats_out, unsigned flags, void* stream) {
  int rc = check_estep_args(c, starts, B, T, stats_out, flags);
  if (rc) return rc;
  if (!c->obs) return fail(SVIHMM_ESTATE, "svihmm_set_series has not been called");
  if (T > c->T_full) return fail(SVIHMM_EINVAL, "T (%d) exceeds the series length (%lld)", T, (long long)c->T_full);
  CU(cudaSetDevice(c->device));
  return estep_impl(c, c->obs, c->obs_dtype, c->mask, starts, B, T, var_x_out, stats_out, flags,
                    (cudaStream_t)stream, trim);
}

extern "C" int svihmm_set_var_init(svihmm_ctx* c, const double* var_init, int loc, void* stream) {
  if (!c) return fail(SVIHMM_EINVAL, "ctx is NULL");
  if (!c->have_globals) return fail(SVIHMM_ESTATE, "svihmm_set_globals has not been called");
  CU(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  c->user_init = var_init != nullptr;
  if (var_init && (rc = copy_in(c->vinit + c->K, var_init, sizeof(double) * c->K, loc, st))) return rc;
  return run_global(c, GM_PREP, nullptr, 0.0, 0.0, 0.0, st);
}

// ---- forward-filter backward-sampling (hmm_fast.pyx:43-124) ---------------------------------------
extern "C" int svihmm_ffbs(svihmm_ctx* c, const double* var_init, int64_t start, int T, int nsamples,
                           uint64_t seed, int32_t* z_out, int loc, void* stream) {
  if (!c || !var_init || !z_out) return fail(SVIHMM_EINVAL, "NULL argument");
  if (!c->have_globals) return fail(SVIHMM_ESTATE, "svihmm_set_globals has not been called");
  if (!c->obs) return fail(SVIHMM_ESTATE, "svihmm_set_series has not been called");
  if (T < 1 || nsamples < 1 || start < 0 || start + T > c->T_full)
    return fail(SVIHMM_EINVAL, "window [%lld, +%d) outside the series / nsamples = %d", (long long)start, T, nsamples);
  CU(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int K = c->K;
  // scratch: P' (K*K) + pi0' (K) floats, var_init (K) doubles, the start index, the sampled paths
  float* Pf = nullptr; double* vi = nullptr; int64_t* dstart = nullptr; int* zd = nullptr;
  CU(dalloc(&Pf, (size_t)K * K + K)); CU(dalloc(&vi, (size_t)K)); CU(dalloc(&dstart, (size_t)1));
  CU(dalloc(&zd, (size_t)nsamples * T));
  int rc = copy_in(vi, var_init, sizeof(double) * K, loc, st);
  if (!rc) rc = copy_in(dstart, &start, sizeof(int64_t), SVIHMM_LOC_HOST, st);
  if (!rc) {
    k_ffbs_prep<<<1, 256, 0, st>>>(K, c->W, vi, Pf, Pf + (size_t)K * K);
    c->launches++;
    // expected log-likelihoods + the scaled forward filter of this ONE window with (P', pi0') swapped in;
    // the marginals / statistics the E-step also produces are discarded (workspace only)
    float* Pt_keep = c->Pt; float* pi0_keep = c->pi0;
    c->Pt = Pf; c->pi0 = Pf + (size_t)K * K;
    rc = estep_impl(c, c->obs, c->obs_dtype, c->mask, dstart, 1, T, nullptr, c->stage_stats,
                    SVIHMM_KEEP_LOCALS, st);
    c->Pt = Pt_keep; c->pi0 = pi0_keep;
  }
  if (!rc) {
    k_ffbs_sample<<<(nsamples * 32 + 127) / 128, 128, 0, st>>>(nsamples, T, K, c->alpha_ws, Pf, seed, zd);
    c->launches++;
    cudaError_t e = cudaMemcpyAsync(z_out, zd, sizeof(int) * (size_t)nsamples * T,
                                    loc == SVIHMM_LOC_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = fail(SVIHMM_ECUDA, "svihmm_ffbs: %s", cudaGetErrorString(e));
  } else cudaStreamSynchronize(st);
  cudaFree(Pf); cudaFree(vi); cudaFree(dstart); cudaFree(zd);
  return rc;
}

// gather B windows of T rows (row = rowbytes) from a mapped host (or device) series into a dense
// [B][T] staging buffer; also the mask bytes.  A FEW long-lived CTAs with several 16-byte loads in
// flight per thread: reads of mapped host memory take microseconds each, and a wide grid of
// short CTAs would hold the SMs' thread slots for that long and lock a concurrently running
// E-step kernel out (measured: fused E-step 65 -> 157 us beside a 1024-CTA gather).
#define GATHER_CTAS 8
#define GATHER_UNROLL 8
static int gather_ctas() {
  static const int n = getenv("SVIHMM_GATHER_CTAS") ? std::max(1, atoi(getenv("SVIHMM_GATHER_CTAS"))) : GATHER_CTAS;
  return n;
}
__global__ void __launch_bounds__(256)
k_gather_windows(int B, int T, int rowbytes, const uint8_t* __restrict__ src, const uint8_t* __restrict__ msrc,
                 const int64_t* __restrict__ starts, uint8_t* __restrict__ dst, uint8_t* __restrict__ mdst,
                 int64_t* __restrict__ dense_starts, int vec16) {
  const size_t wbytes = (size_t)T * rowbytes;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (size_t)gridDim.x * blockDim.x;
  if (vec16) {
    const size_t upw = wbytes / 16, total = upw * B;               // 16-byte units per window / in all
    for (size_t i0 = tid; i0 < total; i0 += nth * GATHER_UNROLL) {
      int4 v[GATHER_UNROLL];
#pragma unroll
      for (int u = 0; u < GATHER_UNROLL; ++u) {
        const size_t i = i0 + (size_t)u * nth;
        if (i < total) {
          const size_t b = i / upw, k = i - b * upw;
          v[u] = *((const int4*)(src + (size_t)starts[b] * rowbytes) + k);
        }
      }
#pragma unroll
      for (int u = 0; u < GATHER_UNROLL; ++u) {
        const size_t i = i0 + (size_t)u * nth;
        if (i < total) ((int4*)dst)[i] = v[u];
      }
    }
  } else {
    const size_t upw = wbytes / 4, total = upw * B;
    for (size_t i = tid; i < total; i += nth) {
      const size_t b = i / upw, k = i - b * upw;
      ((uint32_t*)dst)[i] = *((const uint32_t*)(src + (size_t)starts[b] * rowbytes) + k);
    }
  }
  if (msrc) {
    const size_t total = (size_t)B * T;
    for (size_t i = tid; i < total; i += nth) { const size_t b = i / T; mdst[i] = msrc[starts[b] + (i - b * T)]; }
  }
  for (size_t b = tid; b < (size_t)B; b += nth) dense_starts[b] = (int64_t)b * T;
}
// The same gather on the bulk-copy engine: ONE thread per CTA moves 4 KB pieces host -> shared memory
// (cp.async.bulk over PCIe, completion on an mbarrier) -> staging buffer (cp.async.bulk shared -> global),
// a ring of 4 slots per CTA, 8 CTAs: 128 KB in flight, 46 GB/s stand-alone with 8 threads in all against 36-44 GB/s
// for the load/store kernel with 2048 threads (scripts/probes/tma_gather_probe.cu).  Beside the E-step kernel the
// gather settles at ~39 GB/s whatever the geometry (4.2 MB in 105-108 us); what the geometry decides is how much the
// gather CTAs disturb the placement of the E-step CTAs (two of those fill an SM's shared memory): measured
// end to end at c2, CTAs x slots: 32 x 2 -> 2.22 M E-steps/s (E-step kernel 80 us), 64 x 2 -> 1.63 M (107 us),
// 16 x 2 -> 2.37 M, 8 x 4 -> 2.43 M (66 us).  Lanes 1..31 gather the mask bytes.
#define GB_PIECE 4096
#define GB_SLOTS 4
#define GB_CTAS 8
__global__ void __launch_bounds__(32)
k_gather_bulk(int B, int T, int rowbytes, const uint8_t* __restrict__ src, const uint8_t* __restrict__ msrc,
              const int64_t* __restrict__ starts, uint8_t* __restrict__ dst, uint8_t* __restrict__ mdst,
              int64_t* __restrict__ dense_starts, const int nslots, const int piece) {
  extern __shared__ __align__(128) uint8_t ring[];          // [nslots][piece]
  __shared__ __align__(8) unsigned long long full[GB_SLOTS];
  const size_t wbytes = (size_t)T * rowbytes;
  if (threadIdx.x == 0) {
    for (int s = 0; s < nslots; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(full + s)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const int ppw = (int)((wbytes + piece - 1) / piece);
    const long long np = (long long)B * ppw;
    const int mine = np > (long long)blockIdx.x ? (int)((np - 1 - blockIdx.x) / gridDim.x) + 1 : 0;
    int issued = 0, done = 0;
    while (done < mine) {
      while (issued < mine && issued < done + nslots) {
        // slot reuse: every store issued so far has finished READING its slot
        if (issued >= nslots) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        const long long p = blockIdx.x + (long long)issued * gridDim.x;
        const int b = (int)(p / ppw), k = (int)(p - (long long)b * ppw), s = issued % nslots;
        const unsigned nb = (unsigned)min((size_t)piece, wbytes - (size_t)k * piece);
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(full + s);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(nb) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(ring + s * piece)),
                       "l"(src + (size_t)starts[b] * rowbytes + (size_t)k * piece), "r"(nb), "r"(bar) : "memory");
        ++issued;
      }
      const int s = done % nslots; const unsigned par = (done / nslots) & 1;
      // the piece is microseconds away: sleep between polls (a spinning thread takes issue slots from the E-step CTA
      // that shares the SM, and a persistent kernel is as slow as its slowest CTA)
      for (;;) {
        unsigned ok = 0;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(full + s)), "r"(par) : "memory");
        if (ok) break;
        __nanosleep(250);
      }
      const long long p = blockIdx.x + (long long)done * gridDim.x;
      const int b = (int)(p / ppw), k = (int)(p - (long long)b * ppw);
      const unsigned nb = (unsigned)min((size_t)piece, wbytes - (size_t)k * piece);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                   ::"l"(dst + (size_t)b * wbytes + (size_t)k * piece), "r"((uint32_t)__cvta_generic_to_shared(ring + s * piece)), "r"(nb) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      ++done;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  } else {
    const size_t tid = (size_t)blockIdx.x * 31 + (threadIdx.x - 1), nth = (size_t)gridDim.x * 31;
    if (msrc) {
      const size_t total = (size_t)B * T;
      for (size_t i = tid; i < total; i += nth) { const size_t b = i / T; mdst[i] = msrc[starts[b] + (i - b * T)]; }
    }
    for (size_t b = tid; b < (size_t)B; b += nth) dense_starts[b] = (int64_t)b * T;
  }
}
// windows out of the mapped host series: the bulk-copy engine when rows are 16-byte multiples, else loads/stores
static void launch_gather(int B, int T, int rowbytes, const uint8_t* src, const uint8_t* msrc, const int64_t* starts,
                          uint8_t* dst, uint8_t* mdst, int64_t* dense_starts, cudaStream_t st, bool small_smem) {
  static const bool zc = getenv("SVIHMM_GATHER_ZEROCOPY") != nullptr;    // A/B switch, read once
  static const int env_ctas = getenv("SVIHMM_GATHER_CTAS") ? std::max(1, atoi(getenv("SVIHMM_GATHER_CTAS"))) : 0;
  // beside kernels that fill an SM's shared memory on their own (the persistent tcgen05 kernels: 212-226 KB) only a
  // small CTA can be placed: 2 slots of 2 KB on four times the CTAs; else 4 slots of 4 KB on GB_CTAS CTAs (see above)
  const int nslots = small_smem ? 2 : GB_SLOTS, piece = small_smem ? GB_PIECE / 2 : GB_PIECE;
  const int gb_ctas = env_ctas ? env_ctas : (small_smem ? 4 * GB_CTAS : GB_CTAS);
  const int vec16 = (rowbytes % 16 == 0) && (((uintptr_t)src) % 16 == 0);
  if (vec16 && !zc && ((uintptr_t)dst % 16) == 0)
    k_gather_bulk<<<gb_ctas, 32, (size_t)nslots * piece, st>>>(B, T, rowbytes, src, msrc, starts, dst, mdst, dense_starts, nslots, piece);
  else
    k_gather_windows<<<gather_ctas(), 256, 0, st>>>(B, T, rowbytes, src, msrc, starts, dst, mdst, dense_starts, vec16);
}

extern "C" int svihmm_estep_host(svihmm_ctx* c, const int64_t* starts_host, int B, int T,
                                 float* var_x_host, double* stats_host, unsigned flags, void* stream) {
  int rc = check_estep_args(c, starts_host, B, T, stats_host, flags);
  if (rc) return rc;
  if (!c->hobs) return fail(SVIHMM_ESTATE, "svihmm_set_series_streamed has not been called");
  for (int b = 0; b < B; ++b)
    if (starts_host[b] < 0 || starts_host[b] + T > c->hT_full)
      return fail(SVIHMM_EINVAL, "window %d = [%lld, +%d) outside the series of length %lld", b,
                  (long long)starts_host[b], T, (long long)c->hT_full);
  CU(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t rows = (size_t)B * T, es = esize(c->h_dtype), rowbytes = (size_t)c->OD * es;
  if (rows > c->stage_rows) {
    if (c->stage_obs) CU(cudaFree(c->stage_obs));
    if (c->stage_mask) CU(cudaFree(c->stage_mask));
    c->stage_obs = nullptr; c->stage_mask = nullptr; c->stage_rows = 0;
    CU(cudaMalloc(&c->stage_obs, rows * c->OD * 8));
    CU(cudaMalloc((void**)&c->stage_mask, rows));
    c->stage_rows = rows;
  }
  if ((size_t)B > c->stage_B) {
    if (c->stage_src) CU(cudaFree(c->stage_src));
    if (c->stage_starts) CU(cudaFree(c->stage_starts));
    c->stage_src = nullptr; c->stage_starts = nullptr; c->stage_B = 0;
    CU(dalloc(&c->stage_src, (size_t)B)); CU(dalloc(&c->stage_starts, (size_t)B));
    c->stage_B = B;
  }
  const bool has_mask = c->hmask != nullptr;
  if (c->hobs_dev) {
    // GPU-side gather straight out of the page-locked host series
    CU(cudaMemcpyAsync(c->stage_src, starts_host, sizeof(int64_t) * B, cudaMemcpyHostToDevice, st));
    PhaseTimer pt(c, PH_GATHER, st);
    launch_gather(B, T, (int)rowbytes, (const uint8_t*)c->hobs_dev, has_mask ? c->hmask_dev : nullptr, c->stage_src,
                  (uint8_t*)c->stage_obs, c->stage_mask, c->stage_starts, st, c->K > 32);
    LAUNCHED(c);
  } else {
    // CPU gather into pinned staging, one H2D copy
    if (rows > c->pin_rows) {
      if (c->pin_obs) CU(cudaFreeHost(c->pin_obs));
      if (c->pin_mask) CU(cudaFreeHost(c->pin_mask));
      c->pin_obs = nullptr; c->pin_mask = nullptr; c->pin_rows = 0;
      CU(cudaMallocHost(&c->pin_obs, rows * c->OD * 8 + sizeof(int64_t) * B));
      CU(cudaMallocHost((void**)&c->pin_mask, rows));
      c->pin_rows = rows;
    }
    CU(cudaStreamSynchronize(st));    // the previous step may still be reading the pinned staging
    for (int b = 0; b < B; ++b) {
      memcpy((uint8_t*)c->pin_obs + (size_t)b * T * rowbytes,
             (const uint8_t*)c->hobs + (size_t)starts_host[b] * rowbytes, (size_t)T * rowbytes);
      if (has_mask) memcpy(c->pin_mask + (size_t)b * T, c->hmask + starts_host[b], (size_t)T);
    }
    int64_t* ds = (int64_t*)((uint8_t*)c->pin_obs + rows * c->OD * 8);
    for (int b = 0; b < B; ++b) ds[b] = (int64_t)b * T;
    CU(cudaMemcpyAsync(c->stage_obs, c->pin_obs, rows * rowbytes, cudaMemcpyHostToDevice, st));
    if (has_mask) CU(cudaMemcpyAsync(c->stage_mask, c->pin_mask, rows, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(c->stage_starts, ds, sizeof(int64_t) * B, cudaMemcpyHostToDevice, st));
  }
  if (var_x_host && rows * c->K > c->hostq_cap) {
    if (c->hostq_ws) CU(cudaFree(c->hostq_ws));
    c->hostq_ws = nullptr; c->hostq_cap = 0;
    CU(dalloc(&c->hostq_ws, rows * c->K));
    c->hostq_cap = rows * c->K;
  }
  rc = estep_impl(c, c->stage_obs, c->h_dtype, has_mask ? c->stage_mask : nullptr, c->stage_starts, B, T,
                  var_x_host ? c->hostq_ws : nullptr, c->stage_stats, flags, st);
  if (rc) return rc;
  CU(cudaMemcpyAsync(stats_host, c->stage_stats, sizeof(double) * c->slen, cudaMemcpyDeviceToHost, st));
  if (var_x_host)
    CU(cudaMemcpyAsync(var_x_host, c->hostq_ws, sizeof(float) * rows * c->K, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return SVIHMM_OK;
}

// ---- streamed step: double-buffered window staging ------------------------------------------
static int sg_setup(svihmm_ctx* c) {
  if (c->sg_init) return SVIHMM_OK;
  CU(cudaStreamCreateWithFlags(&c->cstream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->dstream, cudaStreamNonBlocking));
  for (int i = 0; i < SVIHMM_NSLOT; ++i) {
    CU(cudaEventCreateWithFlags(&c->ev_gathered[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_consumed[i], cudaEventDisableTiming));
  }
  CU(cudaEventCreateWithFlags(&c->ev_stats, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_read, cudaEventDisableTiming));
  CU(cudaMallocHost((void**)&c->pin_stats, sizeof(double) * c->slen));
  c->sg_init = 1;
  return SVIHMM_OK;
}

static void sg_teardown(svihmm_ctx* c) {
  if (!c->sg_init) return;
  cudaStreamSynchronize(c->cstream); cudaStreamSynchronize(c->dstream);
  for (int i = 0; i < SVIHMM_NSLOT; ++i) {
    if (c->sg_obs[i]) cudaFree(c->sg_obs[i]);
    if (c->sg_mask[i]) cudaFree(c->sg_mask[i]);
    if (c->sg_src[i]) cudaFree(c->sg_src[i]);
    if (c->sg_dense[i]) cudaFree(c->sg_dense[i]);
    if (c->sg_pin_starts[i]) cudaFreeHost(c->sg_pin_starts[i]);
    if (c->sg_pin_obs[i]) cudaFreeHost(c->sg_pin_obs[i]);
    if (c->sg_pin_mask[i]) cudaFreeHost(c->sg_pin_mask[i]);
    c->sg_pin_obs[i] = nullptr; c->sg_pin_mask[i] = nullptr; c->sg_pin_rows[i] = 0;
    cudaEventDestroy(c->ev_gathered[i]); cudaEventDestroy(c->ev_consumed[i]);
  }
  cudaEventDestroy(c->ev_stats); cudaEventDestroy(c->ev_read);
  cudaFreeHost(c->pin_stats);
  cudaStreamDestroy(c->cstream); cudaStreamDestroy(c->dstream);
  c->sg_init = 0;
}

static int sg_reserve(svihmm_ctx* c, int s, int B, int T) {
  const size_t rows = (size_t)B * T;
  if (rows > c->sg_rows[s]) {
    if (c->sg_obs[s]) CU(cudaFree(c->sg_obs[s]));
    if (c->sg_mask[s]) CU(cudaFree(c->sg_mask[s]));
    c->sg_obs[s] = nullptr; c->sg_mask[s] = nullptr; c->sg_rows[s] = 0;
    CU(cudaMalloc(&c->sg_obs[s], rows * c->OD * 8));
    CU(cudaMalloc((void**)&c->sg_mask[s], rows));
    c->sg_rows[s] = rows;
  }
  if ((size_t)B > c->sg_B[s]) {
    if (c->sg_src[s]) CU(cudaFree(c->sg_src[s]));
    if (c->sg_dense[s]) CU(cudaFree(c->sg_dense[s]));
    if (c->sg_pin_starts[s]) CU(cudaFreeHost(c->sg_pin_starts[s]));
    c->sg_src[s] = nullptr; c->sg_dense[s] = nullptr; c->sg_pin_starts[s] = nullptr; c->sg_B[s] = 0;
    CU(dalloc(&c->sg_src[s], (size_t)B)); CU(dalloc(&c->sg_dense[s], (size_t)B));
    CU(cudaMallocHost((void**)&c->sg_pin_starts[s], sizeof(int64_t) * B));
    c->sg_B[s] = B;
  }
  return SVIHMM_OK;
}

// enqueue on q: windows starts_host[0..B) of the page-locked host series -> staging slot s
static int sg_gather(svihmm_ctx* c, int s, const int64_t* starts_host, int B, int T, cudaStream_t q) {
  const size_t rowbytes = (size_t)c->OD * esize(c->h_dtype);
  const bool has_mask = c->hmask != nullptr;
  memcpy(c->sg_pin_starts[s], starts_host, sizeof(int64_t) * B);
  PhaseTimer pt(c, PH_GATHER, q);
  if (!c->hobs_dev) {
    // the series could not be page-locked (e.g. a memmap larger than host memory): the CPU gathers the
    // windows into pinned staging of this slot (page faults served by the OS), one H2D copy per table
    const size_t rows = (size_t)B * T;
    if (rows > c->sg_pin_rows[s]) {
      if (c->sg_pin_obs[s]) CU(cudaFreeHost(c->sg_pin_obs[s]));
      if (c->sg_pin_mask[s]) CU(cudaFreeHost(c->sg_pin_mask[s]));
      c->sg_pin_obs[s] = nullptr; c->sg_pin_mask[s] = nullptr; c->sg_pin_rows[s] = 0;
      CU(cudaMallocHost(&c->sg_pin_obs[s], rows * rowbytes + sizeof(int64_t) * B));
      CU(cudaMallocHost((void**)&c->sg_pin_mask[s], rows));
      c->sg_pin_rows[s] = rows;
    }
    for (int b = 0; b < B; ++b) {
      memcpy((uint8_t*)c->sg_pin_obs[s] + (size_t)b * T * rowbytes, (const uint8_t*)c->hobs + (size_t)starts_host[b] * rowbytes,
             (size_t)T * rowbytes);
      if (has_mask) memcpy(c->sg_pin_mask[s] + (size_t)b * T, c->hmask + starts_host[b], (size_t)T);
    }
    int64_t* ds = (int64_t*)((uint8_t*)c->sg_pin_obs[s] + rows * rowbytes);
    for (int b = 0; b < B; ++b) ds[b] = (int64_t)b * T;
    CU(cudaMemcpyAsync(c->sg_obs[s], c->sg_pin_obs[s], rows * rowbytes, cudaMemcpyHostToDevice, q));
    if (has_mask) CU(cudaMemcpyAsync(c->sg_mask[s], c->sg_pin_mask[s], rows, cudaMemcpyHostToDevice, q));
    CU(cudaMemcpyAsync(c->sg_dense[s], ds, sizeof(int64_t) * B, cudaMemcpyHostToDevice, q));
    c->sg_valid[s] = 1; c->sg_T[s] = T; c->sg_nB[s] = B;
    return SVIHMM_OK;
  }
  // the GPU gathers the windows itself out of the mapped host series (one small persistent kernel)
  CU(cudaMemcpyAsync(c->sg_src[s], c->sg_pin_starts[s], sizeof(int64_t) * B, cudaMemcpyHostToDevice, q));
  launch_gather(B, T, (int)rowbytes, (const uint8_t*)c->hobs_dev, has_mask ? c->hmask_dev : nullptr, c->sg_src[s],
                (uint8_t*)c->sg_obs[s], c->sg_mask[s], c->sg_dense[s], q, c->K > 32);
  LAUNCHED(c);
  c->sg_valid[s] = 1; c->sg_T[s] = T; c->sg_nB[s] = B;
  return SVIHMM_OK;
}

static int check_windows(svihmm_ctx* c, const int64_t* starts_host, int B, int T) {
  for (int b = 0; b < B; ++b)
    if (starts_host[b] < 0 || starts_host[b] + T > c->hT_full)
      return fail(SVIHMM_EINVAL, "window %d = [%lld, +%d) outside the series of length %lld", b,
                  (long long)starts_host[b], T, (long long)c->hT_full);
  return SVIHMM_OK;
}

static int sg_find(const svihmm_ctx* c, const int64_t* starts_host, int B, int T) {
  for (int i = 0; i < SVIHMM_NSLOT; ++i)
    if (c->sg_valid[i] && c->sg_T[i] == T && c->sg_nB[i] == B &&
        memcmp(c->sg_pin_starts[i], starts_host, sizeof(int64_t) * B) == 0) return i;
  return -1;
}

// replacement: the oldest slot whose minibatch has already been consumed (or that is empty), else
// the oldest staged-but-unused one; never `keep` (the slot of the E-step being enqueued, or -1)
static int sg_victim(const svihmm_ctx* c, int keep) {
  int v = -1;
  for (int pass = 0; pass < 2 && v < 0; ++pass)
    for (int i = 0; i < SVIHMM_NSLOT; ++i) {
      if (i == keep || (pass == 0 && c->sg_pending[i])) continue;
      if (v < 0 || c->sg_age[i] < c->sg_age[v]) v = i;
    }
  return v;
}

static int sg_checks(svihmm_ctx* c, const int64_t* starts_host, int B, int T) {
  if (!c || !starts_host) return fail(SVIHMM_EINVAL, "NULL argument");
  if (B < 1 || T < 1) return fail(SVIHMM_EINVAL, "B (%d) and T (%d) must be >= 1", B, T);
  if (!c->hobs) return fail(SVIHMM_ESTATE, "svihmm_set_series_streamed has not been called");
  return check_windows(c, starts_host, B, T);
}

// gather into slot s on stream q, ordered after the last E-step that read the slot
static int sg_fill(svihmm_ctx* c, int s, const int64_t* starts_host, int B, int T, cudaStream_t q) {
  int rc;
  CU(cudaEventSynchronize(c->ev_gathered[s]));                  // pinned starts of this slot are free again
  c->sg_valid[s] = 0;
  if ((rc = sg_reserve(c, s, B, T))) return rc;
  CU(cudaStreamWaitEvent(q, c->ev_consumed[s], 0));
  if ((rc = sg_gather(c, s, starts_host, B, T, q))) return rc;
  CU(cudaEventRecord(c->ev_gathered[s], q));
  c->sg_age[s] = ++c->sg_clock;
  c->sg_pending[s] = 1;
  return SVIHMM_OK;
}

extern "C" int svihmm_prefetch_windows(svihmm_ctx* c, const int64_t* starts_host, int B, int T) {
  int rc = sg_checks(c, starts_host, B, T);
  if (rc) return rc;
  CU(cudaSetDevice(c->device));
  if ((rc = sg_setup(c))) return rc;
  if (sg_find(c, starts_host, B, T) >= 0) return SVIHMM_OK;      // already staged (or on its way)
  return sg_fill(c, sg_victim(c, -1), starts_host, B, T, c->cstream);
}

extern "C" int svihmm_estep_streamed(svihmm_ctx* c, const int64_t* starts_host, int B, int T,
                                     const int64_t* next_starts_host, float* var_x_dev,
                                     double* stats_dev, unsigned flags, void* stream) {
  int rc = check_estep_args(c, starts_host, B, T, stats_dev, flags);
  if (rc) return rc;
  if ((rc = sg_checks(c, starts_host, B, T))) return rc;
  if (next_starts_host && (rc = check_windows(c, next_starts_host, B, T))) return rc;
  CU(cudaSetDevice(c->device));
  if ((rc = sg_setup(c))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  int s = sg_find(c, starts_host, B, T);
  if (s >= 0) {
    CU(cudaStreamWaitEvent(st, c->ev_gathered[s], 0));          // prefetched on cstream
  } else {
    s = sg_victim(c, -1);
    CU(cudaStreamWaitEvent(st, c->ev_gathered[s], 0));          // a prefetch into this slot may be in flight
    if ((rc = sg_fill(c, s, starts_host, B, T, st))) return rc;
  }
  c->sg_age[s] = ++c->sg_clock;
  c->sg_pending[s] = 0;
  rc = estep_impl(c, c->sg_obs[s], c->h_dtype, c->hmask ? c->sg_mask[s] : nullptr, c->sg_dense[s], B, T,
                  var_x_dev, stats_dev, flags, st);
  if (rc) return rc;
  CU(cudaEventRecord(c->ev_consumed[s], st));
  if (next_starts_host && sg_find(c, next_starts_host, B, T) < 0)
    if ((rc = sg_fill(c, sg_victim(c, s), next_starts_host, B, T, c->cstream))) return rc;
  return SVIHMM_OK;
}

extern "C" int svihmm_svi_step_host(svihmm_ctx* c, const int64_t* starts_host, int B, int T,
                                    const int64_t* next_starts_host, double* stats_host, unsigned flags,
                                    double lrate, double bA, double bE, void* stream) {
  if (!c || !stats_host) return fail(SVIHMM_EINVAL, "NULL argument");
  if (!c->have_prior) return fail(SVIHMM_ESTATE, "globals and priors must be set first");
  int rc = svihmm_estep_streamed(c, starts_host, B, T, next_starts_host, nullptr, c->stage_stats, flags, stream);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaEventRecord(c->ev_stats, st));
  CU(cudaStreamWaitEvent(c->dstream, c->ev_stats, 0));          // read the result back beside the update
  CU(cudaMemcpyAsync(c->pin_stats, c->stage_stats, sizeof(double) * c->slen, cudaMemcpyDeviceToHost, c->dstream));
  CU(cudaEventRecord(c->ev_read, c->dstream));
  if ((rc = run_global(c, GM_SVI, c->stage_stats, lrate, bA, bE, st))) return rc;
  CU(cudaStreamSynchronize(st));
  CU(cudaEventSynchronize(c->ev_read));
  memcpy(stats_host, c->pin_stats, sizeof(double) * c->slen);
  return SVIHMM_OK;
}

extern "C" int svihmm_global_update(svihmm_ctx* c, const double* stats, double lrate, double bA,
                                    double bE, void* stream) {
  if (!c || !stats) return fail(SVIHMM_EINVAL, "NULL argument");
  if (!c->have_globals || !c->have_prior) return fail(SVIHMM_ESTATE, "globals and priors must be set first");
  CU(cudaSetDevice(c->device));
  return run_global(c, GM_SVI, stats, lrate, bA, bE, (cudaStream_t)stream);
}

// ---- multi-GPU: one-shot all-reduce over NVLink peer memory fused into the global step --------------
extern "C" int svihmm_comm_attach(svihmm_ctx* c, int rank, int world, const uint64_t* peer_ptrs) {
  if (!c || !peer_ptrs) return fail(SVIHMM_EINVAL, "NULL argument");
  if (world < 1 || world > 8 || rank < 0 || rank >= world) return fail(SVIHMM_EINVAL, "rank %d / world %d (one node, <= 8 GPUs)", rank, world);
  c->comm_world = world; c->comm_rank = rank; c->comm_seq = 0;
  for (int p = 0; p < world; ++p) c->comm_peer[p] = (void*)(uintptr_t)peer_ptrs[p];
  return SVIHMM_OK;
}

extern "C" size_t svihmm_comm_buffer_len(const svihmm_ctx* c) {
  return c ? (size_t)8 * comm_blocks(c) + (size_t)2 * 8 * c->slen : 0;      // sized for any world <= 8
}

extern "C" int svihmm_global_update_peers(svihmm_ctx* c, const double* stats, double lrate, double bA, double bE,
                                          void* stream) {
  if (!c || !stats) return fail(SVIHMM_EINVAL, "NULL argument");
  if (!c->have_globals || !c->have_prior) return fail(SVIHMM_ESTATE, "globals and priors must be set first");
  if (c->comm_world < 2) return fail(SVIHMM_ESTATE, "svihmm_comm_attach has not been called with world >= 2");
  CU(cudaSetDevice(c->device));
  return run_global(c, GM_SVI, stats, lrate, bA, bE, (cudaStream_t)stream, true);
}

extern "C" int svihmm_get_reduced_stats(svihmm_ctx* c, double* dst, int loc, void* stream) {
  if (!c || !dst) return fail(SVIHMM_EINVAL, "NULL argument");
  CU(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaMemcpyAsync(dst, c->stage_stats, sizeof(double) * c->slen,
                     loc == SVIHMM_LOC_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
  if (loc == SVIHMM_LOC_HOST) CU(cudaStreamSynchronize(st));
  return SVIHMM_OK;
}

/* AdaGrad-like transition step of hmmsgd_metaobs.VBHMM(adagrad=True) (:1036-1040):
 * ada_G += (var_tran-1)^2; var_tran <- (1 - ada_G^-1/4)(var_tran-1) + bfact_A*A*ada_G^-1/4 + 1 (lrate unused);
 * on = 1 (re)initialises ada_G to ones (:183). */
extern "C" int svihmm_set_adagrad(svihmm_ctx* c, int on, void* stream) {
  if (!c) return fail(SVIHMM_EINVAL, "ctx is NULL");
  CU(cudaSetDevice(c->device));
  c->adagrad = on ? 1 : 0;
  if (on) {
    std::vector<double> ones((size_t)c->K * c->K, 1.0);
    CU(cudaMemcpyAsync(c->ada_G, ones.data(), sizeof(double) * ones.size(), cudaMemcpyHostToDevice, (cudaStream_t)stream));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
  }
  return SVIHMM_OK;
}

extern "C" int svihmm_batch_update(svihmm_ctx* c, const double* stats, void* stream) {
  if (!c || !stats) return fail(SVIHMM_EINVAL, "NULL argument");
  if (!c->have_globals || !c->have_prior) return fail(SVIHMM_ESTATE, "globals and priors must be set first");
  CU(cudaSetDevice(c->device));
  c->user_init = 1;    // hmmbatchcd.py:179: var_init becomes an explicit Dirichlet parameter
  return run_global(c, GM_BATCH, stats, 0.0, 0.0, 0.0, (cudaStream_t)stream);
}

extern "C" int svihmm_batchsgd_update(svihmm_ctx* c, const double* stats, double lrate, void* stream) {
  if (!c || !stats) return fail(SVIHMM_EINVAL, "NULL argument");
  if (!c->have_globals || !c->have_prior) return fail(SVIHMM_ESTATE, "globals and priors must be set first");
  CU(cudaSetDevice(c->device));
  c->user_init = 1;    // hmmbatchsgd.py:216: var_init = prior_init + q[0]
  return run_global(c, GM_BSGD, stats, lrate, 1.0, 1.0, (cudaStream_t)stream);
}

extern "C" int svihmm_get_locals(svihmm_ctx* c, double* lliks, float* alpha, double* mx, float* cs,
                                 double* logz, int loc, void* stream) {
  if (!c) return fail(SVIHMM_EINVAL, "ctx is NULL");
  if (c->last_B == 0) return fail(SVIHMM_ESTATE, "no E-step has run yet");
  if (c->last_fused && (lliks || alpha || mx || cs))
    return fail(SVIHMM_ESTATE, "lliks/alpha/mx/cs need the last E-step to run with SVIHMM_KEEP_LOCALS");
  CU(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const cudaMemcpyKind kd = loc == SVIHMM_LOC_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  const size_t n = (size_t)c->last_B * c->last_T * c->K;
  if (lliks) CU(cudaMemcpyAsync(lliks, c->ll_ws, sizeof(double) * n, kd, st));
  if (alpha) CU(cudaMemcpyAsync(alpha, c->alpha_ws, sizeof(float) * n, kd, st));
  const size_t bt = (size_t)c->last_B * c->last_T;
  if (mx) CU(cudaMemcpyAsync(mx, c->mx_ws, sizeof(double) * bt, kd, st));
  if (cs) CU(cudaMemcpyAsync(cs, (const float*)(c->mx_ws + bt), sizeof(float) * bt, kd, st));
  if (logz) CU(cudaMemcpyAsync(logz, c->seq_ws, sizeof(double) * 2 * c->last_B, kd, st));
  if (loc == SVIHMM_LOC_HOST) CU(cudaStreamSynchronize(st));
  return SVIHMM_OK;
}

/* Normalised backward messages of the last KEEP_LOCALS E-step (see include/svihmm.h). */
extern "C" int svihmm_get_locals_beta(svihmm_ctx* c, float* beta, float* sb, int loc, void* stream) {
  if (!c) return fail(SVIHMM_EINVAL, "ctx is NULL");
  if (c->last_B == 0 || c->last_fused || !c->last_beta)
    return fail(SVIHMM_ESTATE, "beta/sb need the last E-step to run with SVIHMM_KEEP_LOCALS");
  CU(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  const cudaMemcpyKind kd = loc == SVIHMM_LOC_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  const size_t bt = (size_t)c->last_B * c->last_T;
  if (beta) CU(cudaMemcpyAsync(beta, c->beta_ws, sizeof(float) * bt * c->K, kd, st));
  if (sb) CU(cudaMemcpyAsync(sb, c->sb_ws, sizeof(float) * bt, kd, st));
  if (loc == SVIHMM_LOC_HOST) CU(cudaStreamSynchronize(st));
  return SVIHMM_OK;
}

/* nsteps global steps of hmmsgd_metaobs.VBHMM.infer (:396-439) in one call: the minibatches are
 * given up front (the samplers do not depend on the globals), every step = E-step over its B windows +
 * natural-gradient update with lrate = (it0 + i + tau)^-kappa (:351); with peers != 0 the statistics
 * are summed over the ranks inside the update kernel (svihmm_global_update_peers). */
extern "C" int svihmm_svi_run(svihmm_ctx* c, const int64_t* starts_all, int nsteps, int B, int T,
                              float* var_x_out, double* stats_out, unsigned flags, double tau, double kappa,
                              int64_t it0, double bA, double bE, int peers, void* stream) {
  int rc = check_estep_args(c, starts_all, B, T, stats_out, flags);
  if (rc) return rc;
  if (nsteps < 1) return fail(SVIHMM_EINVAL, "nsteps = %d", nsteps);
  if (!c->obs) return fail(SVIHMM_ESTATE, "svihmm_set_series has not been called");
  if (!c->have_prior) return fail(SVIHMM_ESTATE, "globals and priors must be set first");
  if (T > c->T_full) return fail(SVIHMM_EINVAL, "T (%d) exceeds the series length (%lld)", T, (long long)c->T_full);
  if (peers && c->comm_world < 2) return fail(SVIHMM_ESTATE, "svihmm_comm_attach has not been called with world >= 2");
  CU(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  size_t fsmem = 0;
  static const bool no_pdl = getenv("SVIHMM_NO_PDL") != nullptr;          // A/B switch, read once
  if (!no_pdl && !c->profiling && !b16_eligible(c, B, flags) && pipe_eligible(c, T, flags, &fsmem)) {
    // Pipelined kernel: the steps are chained with programmatic dependent launch and no memset node.
    // E-step i accumulates into acc[i & 1]; update i consumes it and zeroes acc[(i + 1) & 1] for E-step
    // i + 1 (its only writer, which cannot start its atomics before update i has finished).
    for (int j = 0; j < 2; ++j) {
      if (!c->acc_stats[j]) CU(dalloc(&c->acc_stats[j], c->slen));
      CU(cudaMemsetAsync(c->acc_stats[j], 0, sizeof(double) * c->slen, st));
    }
    c->pdl = 1;
    for (int i = 0; i < nsteps && !rc; ++i) {
      rc = estep_impl(c, c->obs, c->obs_dtype, c->mask, starts_all + (size_t)i * B, B, T, var_x_out,
                      c->acc_stats[i & 1], flags, st);
      const double lrate = pow((double)(it0 + i) + tau, -kappa);
      if (!rc) rc = run_global(c, GM_SVI, c->acc_stats[i & 1], lrate, bA, bE, st, peers != 0, c->acc_stats[(i + 1) & 1]);
    }
    c->pdl = 0;
    if (rc) return rc;
    // the last step's statistics (with peers: this rank's own; the sums are read with svihmm_get_reduced_stats)
    CU(cudaMemcpyAsync(stats_out, c->acc_stats[(nsteps - 1) & 1], sizeof(double) * c->slen, cudaMemcpyDeviceToDevice, st));
    return SVIHMM_OK;
  }
  for (int i = 0; i < nsteps; ++i) {
    if ((rc = estep_impl(c, c->obs, c->obs_dtype, c->mask, starts_all + (size_t)i * B, B, T, var_x_out, stats_out,
                         flags, st))) return rc;
    const double lrate = pow((double)(it0 + i) + tau, -kappa);
    if ((rc = run_global(c, GM_SVI, stats_out, lrate, bA, bE, st, peers != 0))) return rc;
  }
  return SVIHMM_OK;
}

/* Tuning knobs (see include/svihmm.h). */
extern "C" int svihmm_set_tuning(svihmm_ctx* c, int key, int value) {
  if (!c) return fail(SVIHMM_EINVAL, "ctx is NULL");
  switch (key) {
    case SVIHMM_TUNE_B16_MIN_B: c->b16_min_B = value; return SVIHMM_OK;
    case SVIHMM_TUNE_SCAN_MIN_T: c->scan_min_T = value; return SVIHMM_OK;
    case SVIHMM_TUNE_NO_HOSTREG: c->no_hostreg = value ? 1 : 0; return SVIHMM_OK;
    default: return fail(SVIHMM_EINVAL, "unknown tuning key %d", key);
  }
}

/* Global part of the variational lower bound (see include/svihmm.h). */
extern "C" int svihmm_global_bound(svihmm_ctx* c, double* out, int include_init, int loc, void* stream) {
  if (!c || !out) return fail(SVIHMM_EINVAL, "NULL argument");
  if (!c->have_globals || !c->have_prior) return fail(SVIHMM_ESTATE, "globals and priors must be set first");
  if (c->C > 1 && !c->have_mix) return fail(SVIHMM_ESTATE, "svihmm_set_mix_weights has not been called");
  CU(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  BoundArgs a;
  a.K = c->K; a.D = c->D; a.KE = c->KE; a.C = c->C; a.kind = c->kind; a.include_init = include_init ? 1 : 0;
  a.plen = c->plen;
  a.W = c->W; a.vinit = c->vinit; a.emit = c->emit; a.prior_tran = c->prior_tran; a.prior_init = c->prior_init;
  a.prior_emit = c->prior_emit; a.omega = c->omega; a.omega_prior = c->omega_prior;
  a.out = c->rowsum;                                   // scratch double (rewritten by every global step)
  CU(cudaMemsetAsync(a.out, 0, sizeof(double), st));
  const size_t smem = c->kind == SVIHMM_EMIT_NIW_FULL ? (2 * (size_t)c->D * c->D + 2 * (size_t)c->D) * sizeof(double) : 0;
  if (smem > 48 * 1024) CU(cudaFuncSetAttribute(k_global_bound, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_global_bound<<<1 + c->KE + (c->C > 1 ? 1 : 0), 256, smem, st>>>(a);
  LAUNCHED(c);
  CU(cudaMemcpyAsync(out, a.out, sizeof(double), loc == SVIHMM_LOC_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
  if (loc == SVIHMM_LOC_HOST) CU(cudaStreamSynchronize(st));
  return SVIHMM_OK;
}

/* Health of the device-resident parameters (see include/svihmm.h). */
extern "C" int svihmm_check(svihmm_ctx* c, void* stream) {
  if (!c) return fail(SVIHMM_EINVAL, "ctx is NULL");
  CU(cudaSetDevice(c->device));
  cudaStream_t st = (cudaStream_t)stream;
  int flags = 0;
  CU(cudaMemcpyAsync(&flags, c->status_dev, sizeof(int), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (flags & 1) {
    CU(cudaMemsetAsync(c->status_dev, 0, sizeof(int), st));
    return fail(SVIHMM_ESTATE, "an emission scale matrix lost positive definiteness in a global step: the float32 "
                "statistics cancel when |mean| >> spread (centre the series, or use a float64 series with the per-phase kernels)");
  }
  return SVIHMM_OK;
}
