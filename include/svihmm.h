/*
 * svihmm.h -- C ABI of the B200-native SVI-HMM local E-step engine (libsvihmm.so).
 *
 * This is the drop-in boundary for the hot path of dillonalaird/pysvihmm: everything
 * hmmsgd_metaobs.VBHMM.infer does per global step between sampling a minibatch of
 * meta-observations and the natural-gradient update (reference hmmsgd_metaobs.py:405-439),
 * and the batch variant hmmbatchcd.VBHMM.infer (hmmbatchcd.py:135-141).
 *
 * The reference has no FFI for this path: its plugin surface is Python subclassing of
 * hmmbase.VariationalHMMBase (hmmbase.py:34-50; local_update :201, forward_msgs :266,
 * backward_msgs :297 are the documented override points).  The entry points below are what
 * such an override binds with ctypes; each cites the reference code it
 * replaces.  Plain pointers and sizes only; no torch / numpy types.
 *
 * Conventions
 *   - Every function returns 0 on success, a negative SVIHMM_E* code on failure;
 *     svihmm_last_error() returns a thread-local message for the last failure.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All work is
 *     enqueued on it; the *_host entry points synchronise the stream before returning.
 *   - `loc` says where a pointer argument lives: SVIHMM_LOC_DEVICE or SVIHMM_LOC_HOST.
 *   - Matrices are row-major (C order), like the reference's numpy arrays.
 *   - One svihmm_ctx per GPU / host thread; a ctx is not re-entrant (the reference's E-step
 *     mutates `self` and is not re-entrant either, hmmsgd_metaobs.py:487-519).
 *   - There is NO CPU fallback: every entry point that computes needs a CUDA device.
 *
 * Packed layouts (all float64)
 *   emission parameters, per state k, `svihmm_emit_param_len()` doubles:
 *     SVIHMM_EMIT_NIW_FULL : [ mu (D) | sigma (D*D) | kappa | nu ]        Gaussian mu_mf,
 *                            sigma_mf, kappa_mf, nu_mf (pybasicbayes/distributions.py:195-212)
 *     SVIHMM_EMIT_NIW_DIAG : [ mu (D) | sigma (D) | kappa (D) | nu (D) ]  D independent 1-D NIWs
 *     SVIHMM_EMIT_CATEGORICAL : [ alpha_mf (D) ]  Dirichlet over D symbols (Categorical,
 *                            pybasicbayes/distributions.py:1273-1418); the series then has ONE
 *                            column holding the symbol index 0..D-1 (NaN / out of range = missing)
 *   sufficient statistics, `svihmm_stats_len()` doubles:
 *     [ A (K*K) | n (K) | sx (K*D) | sxx (K*D*D, or K*D for DIAG) | q0 (K) | tail (4) ]
 *     A    = sum_b sum_t outer(q[t-1], q[t])         (hmmsgd_metaobs.py:876-878)
 *            (+ B*(prior_tran-1) with SVIHMM_ADD_PRIOR, quirk Q5, :876,881)
 *     n,sx,sxx = sum over unmasked rows of q[t,k]*[1, x, x x^T]   (util.py:73-83)
 *            CATEGORICAL: sx[k][c] = sum_t q[t,k] 1[x_t = c] (hmmsgd_metaobs.py:907-926), no sxx
 *     q0   = sum_b q[b,0,:]                           (hmmbatchcd.py:179)
 *     tail = [ sum_b logZ_b, sum_b Q4_b (hmmsgd_metaobs.py:257-271), B, 0 ]
 */
#ifndef SVIHMM_H_
#define SVIHMM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct svihmm_ctx svihmm_ctx;

enum { SVIHMM_OK = 0, SVIHMM_EINVAL = -1, SVIHMM_ECUDA = -2, SVIHMM_ENOMEM = -3, SVIHMM_ESTATE = -4,
       SVIHMM_EUNSUPPORTED = -5 };
enum { SVIHMM_EMIT_NIW_FULL = 0, SVIHMM_EMIT_NIW_DIAG = 1, SVIHMM_EMIT_CATEGORICAL = 2 };
enum { SVIHMM_F32 = 0, SVIHMM_F64 = 1 };
enum { SVIHMM_LOC_DEVICE = 0, SVIHMM_LOC_HOST = 1 };

/* svihmm_estep flags */
enum {
  SVIHMM_WRAP        = 1u << 0, /* include outer(q[T-1], q[0]) (quirk Q2, hmmsgd_metaobs.py:877-878) */
  SVIHMM_ADD_PRIOR   = 1u << 1, /* add (prior_tran-1) once per window (quirk Q5, :876,881)           */
  SVIHMM_MASK_LL     = 1u << 2, /* masked rows carry no evidence: ll[t,:]=0 (:1167-1168,1176)        */
  SVIHMM_EXACT_XI    = 1u << 3, /* A = sum_t true pairwise posterior instead of outer(q,q) (NOT ref) */
  SVIHMM_KEEP_LOCALS = 1u << 4, /* keep lliks/alpha/cs tables for svihmm_get_locals (unfused kernels) */
  SVIHMM_BF16_DENSE  = 1u << 5  /* 64 < K <= 256 (K % 4 == 0): the K x K step of the recursions as a dense
                                   (128 windows x K).(K x K) contraction on tcgen05 tensor cores, bf16
                                   messages with float32 accumulators (BASELINE config 4), and the
                                   statistics as tcgen05 contractions over the windows from bf16-rounded
                                   marginals; marginals and statistics then agree with the float64
                                   reference to ~1e-2 instead of 1e-5 (NOT ref) */
};

const char* svihmm_last_error(void);
int svihmm_version(void);

/* Lifetime.  Replaces VariationalHMMBase.__init__ state (hmmbase.py:67-136). */
int svihmm_create(svihmm_ctx** out, int device, int K, int D, int emission_kind);
int svihmm_destroy(svihmm_ctx* ctx);
size_t svihmm_emit_param_len(const svihmm_ctx* ctx); /* doubles per state */
size_t svihmm_stats_len(const svihmm_ctx* ctx);      /* doubles           */

/* EXTENSION (BASELINE config 5, no mean-field counterpart in the reference: MixtureDistribution,
 * pybasicbayes/models.py:256-300, is Gibbs/EM only): every state emits from a mixture of C NIW
 * components (kind = SVIHMM_EMIT_NIW_FULL or _DIAG) with Dirichlet weights.  Component arrays
 * (svihmm_set_prior / svihmm_set_globals `emit`) then hold K*C rows of svihmm_emit_param_len()
 * doubles, row k*C + c = component c of state k; the statistics are
 *   [ A (K*K) | n (K*C) | sx (K*C*D) | sxx (K*C*DD) | q0 (K) | tail (4) ]
 * with component weights q[t,k] * r[t,k,c], r = softmax_c(E[ln pi_kc] + E[ln N_kc(x_t)])
 * (pybasicbayes/internals/labels.py:52-65), and ll[t,k] = logsumexp_c of the same terms.
 * svihmm_set_mix_weights must be called (with the prior) before the first E-step; omega, omega_prior:
 * K*C Dirichlet parameters.  Only the SVI update (svihmm_global_update) is defined for mixtures. */
int svihmm_create_mix(svihmm_ctx** out, int device, int K, int D, int emission_kind, int C);
int svihmm_set_mix_weights(svihmm_ctx* ctx, const double* omega, const double* omega_prior, int loc,
                           void* stream);
int svihmm_get_mix_weights(svihmm_ctx* ctx, double* omega, int loc, void* stream);

/* Observation series obs (T_full x D, dtype f32/f64) and optional mask (T_full bytes, 1 = missing,
 * hmmbase.py:60-65).  Replaces set_data (hmmbase.py:138-143).
 * loc = DEVICE: pointers are borrowed (caller keeps them alive).  loc = HOST: copied into HBM. */
int svihmm_set_series(svihmm_ctx* ctx, const void* obs, int64_t T_full, int dtype,
                      const uint8_t* mask, int loc, void* stream);

/* Host-resident series for streaming: the buffer is page-locked and mapped so that each step's
 * windows are gathered over PCIe/NVLink-C2C by the GPU itself (gen_synthetic.read_data_mmap
 * feeder, gen_synthetic.py:188-191).  Used by svihmm_estep_host. */
int svihmm_set_series_streamed(svihmm_ctx* ctx, const void* obs_host, int64_t T_full, int dtype,
                               const uint8_t* mask_host);

/* Priors: prior_tran (K*K), prior_init (K, may be NULL = ones), prior_emit (K*param_len).
 * hmmbase.py:102-104. */
int svihmm_set_prior(svihmm_ctx* ctx, const double* prior_tran, const double* prior_init,
                     const double* prior_emit, int loc, void* stream);

/* Global variational parameters: var_tran (K*K Dirichlet), var_init (K, or NULL = recompute the
 * |top eigenvector| of the mean transition matrix each time var_tran changes, quirk Q3,
 * hmmsgd_metaobs.py:413-418), emit (K*param_len).  Derives on device mod_init/mod_tran
 * (hmmsgd_metaobs.py:502-504) and the Cholesky-based emission constants
 * (pybasicbayes/distributions.py:351-366). */
int svihmm_set_globals(svihmm_ctx* ctx, const double* var_tran, const double* var_init,
                       const double* emit, int loc, void* stream);
int svihmm_get_globals(svihmm_ctx* ctx, double* var_tran, double* var_init, double* emit,
                       int loc, void* stream);

/* The E-step for a minibatch of B windows obs[starts[b] : starts[b]+T] of the resident series.
 * Replaces, for all B meta-observations at once, local_update (hmmsgd_metaobs.py:487-519:
 * lliks :508-509, forward_msgs :775-803, backward_msgs :828-855, marginals :516-519),
 * intermediate_pars (:857-904), the accumulation :430-433 and local_lower_bound (:257-271).
 *   starts    : B int64 (device)
 *   var_x_out : B*T*K float32 posterior marginals (device), or NULL
 *   stats_out : svihmm_stats_len() doubles (device)                                         */
int svihmm_estep(svihmm_ctx* ctx, const int64_t* starts, int B, int T, float* var_x_out,
                 double* stats_out, unsigned flags, void* stream);

/* Buffered meta-observations (growBuffer: select_buffer hmmsgd_metaobs.py:579-661 +
 * intermediate_pars_buffer :932-1008): the E-step runs on windows of T = 2*bufferL+1 rows, but only
 * the inner T - 2*trim rows (trim = bufferL - L) feed the statistics, with the wrap-around pair
 * taken inside the inner slice exactly as :957-958 does on var_x[bufferL-L : bufferL+L+1].
 * var_x_out holds the marginals of the whole buffered windows; the log-normaliser tail is that of
 * the whole windows (local_lower_bound reads the full lalpha).  trim = 0 is svihmm_estep. */
int svihmm_estep_buffered(svihmm_ctx* ctx, const int64_t* starts, int B, int T, int trim,
                          float* var_x_out, double* stats_out, unsigned flags, void* stream);

/* Replace the initial-state Dirichlet parameter in use (var_init != NULL), or go back to the
 * stationary vector of the current transition parameters (NULL), and refresh the per-step
 * constants.  select_L / select_buffer / get_local_messages (hmmsgd_metaobs.py:521-700) run with
 * whatever self.var_init holds at that moment, which is the vector of the PREVIOUS minibatch. */
int svihmm_set_var_init(svihmm_ctx* ctx, const double* var_init, int loc, void* stream);

/* Same with HOST buffers (the reference-facing call): windows are gathered from the streamed host
 * series (svihmm_set_series_streamed) host->device, the E-step runs, stats (and var_x if not
 * NULL) are copied device->host, and the stream is synchronised. */
int svihmm_estep_host(svihmm_ctx* ctx, const int64_t* starts_host, int B, int T,
                      float* var_x_host, double* stats_host, unsigned flags, void* stream);

/* Streamed step for training loops over a HOST-resident series (svihmm_set_series_streamed).
 * The sampler does not depend on the global parameters (hmmsgd_metaobs.py:396, metaobs_unif
 * :210-227), so upcoming minibatches can be drawn early and their windows moved host->device
 * while the current minibatch is being processed:
 *   svihmm_prefetch_windows  enqueues, on an internal copy stream, the gather of the windows
 *                            obs[starts[b] : starts[b]+T] into one of SVIHMM ring slots (3); no-op
 *                            if that minibatch is already staged.  Never blocks on compute.
 *   svihmm_estep_streamed    enqueues (no synchronisation) the E-step of a minibatch on `stream`:
 *                            from its staged copy if (starts, B, T) match a prefetched minibatch,
 *                            else after gathering it on `stream`.  Statistics stay in the caller's
 *                            DEVICE buffer so that an all-reduce and svihmm_global_update can follow
 *                            on the same stream.  next_starts_host != NULL is a convenience for
 *                            svihmm_prefetch_windows(next_starts_host, B, T) after the launch.
 *   var_x_dev : B*T*K float32 (device) or NULL;  stats_dev : svihmm_stats_len() doubles (device). */
int svihmm_prefetch_windows(svihmm_ctx* ctx, const int64_t* starts_host, int B, int T);
int svihmm_estep_streamed(svihmm_ctx* ctx, const int64_t* starts_host, int B, int T,
                          const int64_t* next_starts_host, float* var_x_dev, double* stats_dev,
                          unsigned flags, void* stream);

/* One whole global step of hmmsgd_metaobs.VBHMM.infer (:396-439) for one process, HOST buffers in
 * and out: svihmm_estep_streamed + svihmm_global_update on the device-resident statistics, with the
 * minibatch statistics (local_lower_bound terms included) copied to stats_host beside the update;
 * synchronises `stream` before returning. */
int svihmm_svi_step_host(svihmm_ctx* ctx, const int64_t* starts_host, int B, int T,
                         const int64_t* next_starts_host, double* stats_host, unsigned flags,
                         double lrate, double bfact_A, double bfact_E, void* stream);

/* Stochastic natural-gradient step on the resident globals from (all-reduced) statistics:
 * hmmsgd_metaobs.py:1010-1069 with util.py:28-60.  stats on device.
 *   var_tran <- (1-lrate)(var_tran-1) + lrate*bfact_A*A + 1
 *   eta_k    <- (1-lrate) eta_k + lrate (eta_prior + bfact_E * e_k)                          */
int svihmm_global_update(svihmm_ctx* ctx, const double* stats, double lrate, double bfact_A,
                         double bfact_E, void* stream);

/* Multi-GPU (one process per GPU, windows of a minibatch sharded over the ranks): the sum over ranks
 * of the minibatch statistics (the accumulation hmmsgd_metaobs.py:430-433) is taken INSIDE the
 * global-step kernel over NVLink peer memory instead of a separate NCCL all-reduce: every block
 * pushes the statistics it consumes into receive slots of all peers (P2P stores), raises per-block
 * sequence flags, and sums the received copies in rank order (bitwise identical on every rank).
 *   svihmm_comm_buffer_len      doubles each rank must allocate as PEER-ACCESSIBLE, zero-initialised
 *                               device memory (e.g. torch.distributed._symmetric_memory)
 *   svihmm_comm_attach          peer_ptrs[p] = device address of rank p's area as seen from this rank
 *   svihmm_global_update_peers  every rank calls it once per step with ITS statistics (device): the
 *                               exchange and the update hmmsgd_metaobs.py:1010-1084 in one launch
 *   svihmm_get_reduced_stats    the all-reduced statistics of the last step (for the bound / logging) */
size_t svihmm_comm_buffer_len(const svihmm_ctx* ctx);
int svihmm_comm_attach(svihmm_ctx* ctx, int rank, int world, const uint64_t* peer_ptrs);
int svihmm_global_update_peers(svihmm_ctx* ctx, const double* stats, double lrate, double bfact_A,
                               double bfact_E, void* stream);
int svihmm_get_reduced_stats(svihmm_ctx* ctx, double* dst, int loc, void* stream);

/* AdaGrad-like variant of the transition step in svihmm_global_update (hmmsgd_metaobs.py:1036-1040,
 * VBHMM(adagrad=True)): ada_G += (var_tran-1)^2, step size ada_G^(-1/4) per entry instead of lrate;
 * on = 1 (re)initialises ada_G to ones (:183), on = 0 restores the plain step. */
int svihmm_set_adagrad(svihmm_ctx* ctx, int on, void* stream);

/* Batch coordinate-ascent step, hmmbatchcd.py:172-189 + distributions.py:240-276,324-329:
 * var_init = prior_init + q0, var_tran = prior_tran + A, conjugate NIW update per state.
 * stats must come from svihmm_estep with B = 1, flags without WRAP / ADD_PRIOR. */
int svihmm_batch_update(svihmm_ctx* ctx, const double* stats, void* stream);

/* Batch natural-gradient step, hmmbatchsgd.py:202-259: var_init = prior_init + q0,
 *   var_tran <- (1-lrate)(var_tran-1) + lrate*(prior_tran + A - 1) + 1,
 *   eta_k    <- (1-lrate) eta_k + lrate * eta(conjugate posterior of state k)  [= eta_prior + e_k]
 * stats must come from svihmm_estep with B = 1 and SVIHMM_ADD_PRIOR (| SVIHMM_MASK_LL as
 * hmmbatchsgd.infer NaN-masks the observations, :148-149), no SVIHMM_WRAP. */
int svihmm_batchsgd_update(svihmm_ctx* ctx, const double* stats, double lrate, void* stream);

/* Local tables of the last svihmm_estep (valid until the next one; device -> dst at loc):
 *   lliks  B*T*K float64  expected log-likelihoods (self.lliks, hmmsgd_metaobs.py:508-509)
 *   alpha  B*T*K float32  normalised forward messages = softmax_k(self.lalpha[t])
 *   mx     B*T   float64  max_k lliks[t,k]
 *   cs     B*T   float32  forward scale factors c_t, so that
 *                         self.lalpha[t,k] = log alpha[t,k] + sum_{u<=t} (log cs[u] + mx[u])
 *   logz   B*2   float64  per window [logZ, Q4 bound (hmmsgd_metaobs.py:257-271)]
 * Any of them may be NULL.  lliks/alpha/mx/cs need the last E-step to have run with
 * SVIHMM_KEEP_LOCALS (the fused single-kernel path keeps these tables on chip only). */
int svihmm_get_locals(svihmm_ctx* ctx, double* lliks, float* alpha, double* mx, float* cs,
                      double* logz, int loc, void* stream);

/* nsteps global steps of hmmsgd_metaobs.VBHMM.infer (hmmsgd_metaobs.py:396-439) enqueued by ONE call:
 * starts_all is nsteps*B window starts (device; the samplers :210-255 do not depend on the globals, so
 * the minibatches of all steps are drawn up front); step i runs the E-step over its B windows of the
 * resident series and the natural-gradient update with lrate = (it0 + i + tau)^-kappa (:351).
 * var_x_out (B*T*K, may be NULL) and stats_out hold the LAST step's values on return.  peers != 0: the
 * statistics are summed over the attached ranks inside the update (svihmm_global_update_peers). */
int svihmm_svi_run(svihmm_ctx* ctx, const int64_t* starts_all, int nsteps, int B, int T, float* var_x_out,
                   double* stats_out, unsigned flags, double tau, double kappa, int64_t it0, double bfact_A,
                   double bfact_E, int peers, void* stream);

/* Global part of the variational lower bound from the device-resident parameters, one float64 at `loc`:
 * hmmsgd_metaobs.VBHMM.global_lower_bound (hmmsgd_metaobs.py:273-296) = Dirichlet energy + entropy of the
 * transition rows + sum over states of var_emit[k].get_vlb() (Gaussian: pybasicbayes/distributions.py:331-349,
 * Categorical: :1372-1381; per dimension for the diagonal model; mixture weights as Categoricals);
 * include_init != 0 adds the Dirichlet terms of the initial distribution (hmmbase.lower_bound,
 * hmmbase.py:145-199).  The data term is the statistics tail [sum logZ, sum Q4] of the E-step. */
int svihmm_global_bound(svihmm_ctx* ctx, double* out, int include_init, int loc, void* stream);

/* Health check of the device-resident parameters; synchronises.  SVIHMM_ESTATE (and the flag is cleared)
 * if a global step met a non-positive pivot / variance in an emission scale: the emission statistics are
 * accumulated from float32 products (the reference's util.NIW_suffstats is float64), so for a series whose
 * mean is far from zero in units of its spread (|mean| / std >~ 1e3) Sigma = eta3 - kappa mu mu^T cancels.
 * Centre such a series before handing it over. */
int svihmm_check(svihmm_ctx* ctx, void* stream);

/* Tuning knobs.  SVIHMM_TUNE_B16_MIN_B: smallest minibatch (windows per call) that takes the batched
 * tensor-core path for K <= 16 diagonal models (sixteen windows per chain warp, batch16.cuh); smaller
 * calls use the one-CTA-per-window pipelined kernel (default 4096: the measured crossover at the c2
 * shape).  0 disables the batched path, 1 forces it for every eligible call. */
enum { SVIHMM_TUNE_B16_MIN_B = 1,
       /* shortest window that takes the block-parallel scan (K <= 16, few long chains; scan16.cuh)
        * instead of the sequential per-phase recursions; default 4096, 0 = never */
       SVIHMM_TUNE_SCAN_MIN_T = 2,
       /* != 0: svihmm_set_series_streamed does not page-lock the host series; the windows of every step are
        * gathered by the CPU into pinned staging instead (for memory-mapped series larger than host memory,
        * gen_synthetic.read_data_mmap, gen_synthetic.py:188-191).  Set before svihmm_set_series_streamed. */
       SVIHMM_TUNE_NO_HOSTREG = 3 };
int svihmm_set_tuning(svihmm_ctx* ctx, int key, int value);

/* Backward table of the last SVIHMM_KEEP_LOCALS E-step (self.lbeta, hmmsgd_metaobs.py:828-855 /
 * hmmbase.py:297-320):
 *   beta  B*T*K float32  normalised backward messages = softmax_k(self.lbeta[t]); beta[T-1] = 1
 *   sb    B*T   float32  backward scale factors d_t (d_{T-1} = 1), so that
 *                        self.lbeta[t,k] = log beta[t,k] + sum_{u=t}^{T-2} (log sb[u] + mx[u+1])
 * Finite wherever the reference's table is (no reconstruction through the marginals). */
int svihmm_get_locals_beta(svihmm_ctx* ctx, float* beta, float* sb, int loc, void* stream);

/* Forward-filter backward-sampling of state paths for the window obs[start : start+T] (the reference's
 * native kernel hmm_fast.FFBS, hmm_fast.pyx:43-124, bound at hmmbase.py:410-411): forward filter with
 * initial weights psi(var_init+eps) - psi(sum+eps) and transition weights log(var_tran + eps) (:82-100),
 * then nsamples independent backward passes (:103-122) with a Philox stream per sample (the reference
 * draws from libc rand(), so paths agree in distribution only).  z_out: nsamples*T int32 at `loc`.
 * The forward table of the call stays readable through svihmm_get_locals (alpha, cs, mx). */
int svihmm_ffbs(svihmm_ctx* ctx, const double* var_init, int64_t start, int T, int nsamples, uint64_t seed,
                int32_t* z_out, int loc, void* stream);

/* Number of kernels the engine launched on this ctx since creation (bench bookkeeping). */
int64_t svihmm_launch_count(const svihmm_ctx* ctx);

/* Optional per-phase device timing (bench bookkeeping; the reference's only instrumentation is
 * the wall clock iter_time, hmmsgd_metaobs.py:348,441).  While enabled every compute entry point
 * brackets its phases with CUDA events on the caller's stream.  svihmm_get_phase_ms waits for the
 * recorded events, returns the summed milliseconds and the number of recorded intervals per phase
 * since the last call, and resets.  Phases are named by svihmm_phase_name(i), i < SVIHMM_N_PHASES. */
enum { SVIHMM_N_PHASES = 8 };
int svihmm_set_profiling(svihmm_ctx* ctx, int on);
int svihmm_get_phase_ms(svihmm_ctx* ctx, double* ms, int64_t* counts);
const char* svihmm_phase_name(int phase);

#ifdef __cplusplus
}
#endif
#endif /* SVIHMM_H_ */
