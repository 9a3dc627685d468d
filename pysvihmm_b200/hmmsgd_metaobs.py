"""hmmsgd_metaobs.VBHMM over the CUDA engine: SVI with minibatches of meta-observations.

Mirrors reference hmmsgd_metaobs.py: same constructor (:78-208), samplers (:210-255), infer
control flow (:298-485), method names and attributes.  The sequential per-meta-observation loop
(:405-436) becomes ONE batched E-step on the GPU, the accumulation (:430-433) happens inside the
statistics kernel (followed by one NCCL all-reduce when torch.distributed is initialised) and
global_update (:1010-1069) runs on the device-resident parameters.
"""
import sys
import time

import numpy as np
import numpy.random as npr

from . import _lib as L
from .hmmbase import VariationalHMMBase
from .sharding import allreduce_stats, broadcast_minibatch, dist_or_none as _dist, shard_starts

eps = 1e-9
tau0 = 1.
kappa0 = 0.7
metaobs_half0 = 1
mb_sz0 = 1


class MetaObs(object):
    """Inclusive index pair (i1, i2), hmmsgd_metaobs.py:42-45."""

    def __init__(self, i1, i2):
        self.i1 = i1
        self.i2 = i2


class VBHMM(VariationalHMMBase):
    """Stochastic variational inference for finite HMMs with natural-gradient global updates;
    consecutive groups of nodes are sampled as a "meta-observation" (hmmsgd_metaobs.py:47)."""

    @staticmethod
    def make_param_dict(prior_init, prior_tran, prior_emit, tau=tau0, kappa=kappa0,
                        metaobs_half=metaobs_half0, mb_sz=mb_sz0, mask=None):
        """hmmsgd_metaobs.py:59-68."""
        return {'prior_init': prior_init, 'prior_tran': prior_tran, 'prior_emit': prior_emit,
                'mask': mask, 'tau': tau, 'kappa': kappa, 'metaobs_half': metaobs_half, 'mb_sz': mb_sz}

    def set_metaobs_fun(self):
        """hmmsgd_metaobs.py:70-76."""
        if self.metaobs_fun_name == 'unif':
            self.metaobs_fun = self.metaobs_unif
        elif self.metaobs_fun_name == 'noverlap':
            self.metaobs_fun = self.metaobs_noverlap
        else:
            raise RuntimeError("Unknown value for metaobs_fun: %s" % (self.metaobs_fun_name,))

    def __init__(self, obs, prior_init, prior_tran, prior_emit, tau=tau0, kappa=kappa0,
                 metaobs_half=metaobs_half0, mb_sz=mb_sz0, mask=None, full_predprob=False,
                 init_init=None, init_tran=None, maxit=100, verbose=False, adagrad=False,
                 metaobs_fun='unif', seed=None, sts=None, fullpred_freq=10, fullpred_sched=None,
                 growBuffer=False, bufferBudget=False, obs_dtype="f64", device=None,
                 track_elbo=True, pairwise_mode="ref_outer", peer_allreduce=True):
        """hmmsgd_metaobs.py:78-208.  Engine kwargs: obs_dtype, device, track_elbo (False skips the
        per-iteration host read of the bound), pairwise_mode ('ref_outer' = the reference's
        product-of-marginals statistic, 'exact_xi' = true pairwise posteriors)."""
        np.random.seed(seed)
        self.seed = seed
        super(VBHMM, self).__init__(obs, prior_init, prior_tran, prior_emit, mask=mask,
                                    init_init=init_init, init_tran=init_tran, verbose=verbose, sts=sts,
                                    obs_dtype=obs_dtype, device=device)
        self._explicit_init = False     # var_init is recomputed from var_tran (:413-418)
        self.elbo = -np.inf
        self.tau = tau
        self.kappa = kappa
        self.lrate = tau ** (-kappa)
        self.full_predprob = full_predprob
        self.fullpred_freq = fullpred_freq
        self.fullpred_sched = fullpred_sched if fullpred_sched is not None else np.arange(0, maxit, 10)
        self.metaobs_fun_name = metaobs_fun
        self.set_metaobs_fun()
        self.adagrad = adagrad
        if adagrad:
            self.ada_G = 1.0 * np.ones(self.prior_tran.shape)      # :183 (device copy lives in the engine)
        self.maxit = maxit
        self.growBuffer = growBuffer
        self.bufferBudget = bufferBudget
        if metaobs_half is not None and metaobs_half < 1:
            raise RuntimeError("metaobs (%d) must be >= 1." % (metaobs_half,))
        self.metaobs_half = metaobs_half
        self._ctor_half = metaobs_half
        self.mb_sz = mb_sz
        self.cur_mo = None
        self.batchfactor = 1.
        self.track_elbo = track_elbo
        self.peer_allreduce = peer_allreduce   # N > 1: sum over ranks inside the global-step kernel (NVLink P2P)
        self._px = None
        if pairwise_mode not in ("ref_outer", "exact_xi"):
            raise RuntimeError("pairwise_mode must be 'ref_outer' or 'exact_xi'")
        self.pairwise_mode = pairwise_mode
        metaobs_sz = 2 * (metaobs_half if metaobs_half is not None else 1) + 1
        self.var_x = np.random.rand(metaobs_sz, self.K)       # :202 (keeps the RNG stream aligned)
        self.var_x /= np.sum(self.var_x, axis=1)[:, np.newaxis]
        self.lalpha = np.empty((metaobs_sz, self.K))
        self.lbeta = np.empty((metaobs_sz, self.K))
        self.lliks = np.empty((metaobs_sz, self.K))

    # ------------------------------------------------------------------ samplers
    def metaobs_unif(self, N, L, n):
        """hmmsgd_metaobs.py:210-227."""
        c_vec = npr.randint(L, N - 1 - L + 1, n)
        return [MetaObs(c - L, c + L) for c in c_vec]

    def metaobs_noverlap(self, N, L, n):
        """hmmsgd_metaobs.py:229-255 (returns n+1 windows exactly like the reference)."""
        ll, uu = L, N - 1 - L
        c_vec = np.inf * np.ones(n)
        minibatch = list()
        c = npr.randint(ll, uu + 1, 1)[0]
        minibatch.append(MetaObs(c - L, c + L))
        for i in range(n):
            c = npr.randint(ll, uu + 1, 1)[0]
            while np.any(np.abs(c_vec - c) <= L):
                c = npr.randint(ll, uu + 1, 1)[0]
            c_vec[i] = c
            minibatch.append(MetaObs(c - L, c + L))
        return minibatch

    # ------------------------------------------------------------------ bounds
    def local_lower_bound(self):
        """hmmsgd_metaobs.py:257-271 for the last minibatch (sum over its windows; quirk Q4)."""
        return float(self._last_stats_host["lb_q4"])

    def global_lower_bound(self):
        """hmmsgd_metaobs.py:273-296, evaluated on the device from the resident parameters
        (svihmm_global_bound): no per-iteration read-back of the globals."""
        return self._ensure_engine().global_bound(include_init=False)

    def _global_lower_bound_host(self):
        """The same from the host copies of the parameters (the classes' get_vlb); kept as the
        cross-check of the device kernel."""
        self._pull_globals()
        out = self._dirichlet_bound(self.prior_tran, self.var_tran)
        for k in range(self.K):
            out += self.var_emit[k].get_vlb()
        return out

    # ------------------------------------------------------------------ the hot path
    def _flags(self):
        f = L.WRAP | L.ADD_PRIOR
        if self.pairwise_mode == "exact_xi":
            f |= L.EXACT_XI
        return f

    def minibatch_estep(self, minibatch, want_var_x=True, trim=0):
        """Batched replacement of the loop hmmsgd_metaobs.py:405-436 for this rank's share of the
        minibatch; returns the packed, all-reduced statistics tensor (device).  trim > 0: buffered
        meta-observations, statistics from the inner rows only (intermediate_pars_buffer :932-1008)."""
        eng = self._ensure_engine()
        starts = np.array([m.i1 for m in minibatch], dtype=np.int64)
        T = int(minibatch[0].i2 - minibatch[0].i1 + 1)
        self._check_windows(starts, T)
        dist = _dist()
        if dist is not None:
            # every rank must work on the SAME minibatch and window length (a rank-local RNG draw in
            # select_L / the sampler would silently de-synchronise the replicas): rank 0's wins
            starts, T = broadcast_minibatch(starts, T, dist, eng.device)
            starts = shard_starts(starts, dist.get_rank(), dist.get_world_size())
        if len(starts) == 0:
            # more ranks than windows (e.g. mb_sz = 1): this rank contributes zero statistics but
            # still takes part in the sum over ranks below
            vx, stats = None, eng.new_stats().zero_()
        else:
            vx, stats = eng.estep(starts, T, flags=self._flags(), want_var_x=want_var_x, trim=trim)
        if dist is not None and self.peer_allreduce and dist.get_backend() == "nccl":
            if self._px is None:            # the sum over ranks happens inside global_update (P2P over NVLink)
                from .sharding import PeerExchange
                self._px = PeerExchange(eng, dist)   # raises if the ranks cannot address each other's memory;
                                                     # construct with peer_allreduce=False for the NCCL all-reduce
        else:
            allreduce_stats(stats, dist)    # one sum all-reduce of the packed statistics per step
        self._var_x_batch = vx
        self._last_B, self._last_T = len(starts), T
        self._last_B_global = len(minibatch)
        return stats

    def _check_windows(self, starts, T):
        """Window starts handed to the device are validated on the host (the kernels index the
        resident series without bounds checks)."""
        starts = np.asarray(starts, dtype=np.int64)
        if T < 1 or T > self.T or (starts.size and (starts.min() < 0 or starts.max() + T > self.T)):
            raise RuntimeError("meta-observation outside the series: starts in [%d, %d], length %d, T = %d" % (
                int(starts.min()) if starts.size else 0, int(starts.max()) if starts.size else 0, T, self.T))

    def infer(self, adaptive=False, perIter=10, epsilon=1e-6, minHalfL=1, avgResidual=False,
              Lincrement=1, Lcutoff=1000):
        """hmmsgd_metaobs.py:298-485."""
        np.random.seed(self.seed)
        growBuffer, bufferBudget = self.growBuffer, self.bufferBudget
        maxit = self.maxit
        if self.metaobs_fun is None:
            self.set_metaobs_fun()
        self.elbo_vec = np.inf * np.ones(maxit)
        self.iter_time = np.inf * np.ones(maxit)
        mb_sz, Lh = self.mb_sz, self.metaobs_half
        miniL, bufferL = Lh, None
        if (Lh is None or adaptive) and growBuffer:
            raise RuntimeError("Cannot specify both adaptive and buffer simultaneously!")   # :344
        eng = self._ensure_engine()
        if self.adagrad and not getattr(self, "_ada_started", False):
            eng.set_adagrad(True)          # ada_G = ones once (constructor, :183); kept across infer() calls
            self._ada_started = True
        track_init = adaptive or Lh is None or growBuffer       # select_* read the previous var_init
        for it in range(maxit):
            start_time = time.time()
            self.lrate = (it + self.tau) ** (-self.kappa)                 # :351
            if Lh is None or (adaptive and it % perIter == 0):            # :354-369
                Lh = self.select_L(mb_sz, epsilon=epsilon, minHalfL=minHalfL, avgResidual=avgResidual,
                                   Lincrement=Lincrement, Lcutoff=Lcutoff)
                # L stays local as in the reference: global_update keeps scaling with the constructor's
                # metaobs_half (:340,355,1021).  metaobs_half=None (an extension: the reference's
                # constructor raises for it) has no such value, so there the selected L is used.
                if self._ctor_half is None:
                    self.metaobs_half = Lh
                self._resize_locals(2 * Lh + 1)
                miniL = Lh
            if growBuffer and it % perIter == 0:                          # :372-393
                bufferL = self.select_buffer(self.mb_sz, epsilon=epsilon, halfL=Lh, avgResidual=avgResidual,
                                             Lincrement=Lincrement, Lcutoff=Lcutoff)
                self._resize_locals(2 * bufferL + 1)
                miniL = bufferL
                if bufferBudget:
                    mb_sz = self.buffer_budget(bufferL)
            minibatch = self.metaobs_fun(self.T, miniL, mb_sz)            # :396
            self.cur_mo = minibatch[-1]
            stats = self.minibatch_estep(minibatch, want_var_x=False,
                                         trim=(bufferL - Lh) if growBuffer else 0)   # :405-436
            if track_init:
                self.var_init = eng.get_globals()[1]                      # :418, read by the next select_*
            self.global_update(stats)                                     # :439
            if self.track_elbo:
                self._last_stats_host = eng.unpack_stats(stats)
                self.iter_time[it] = time.time() - start_time
                lb = self.local_lower_bound() + self.global_lower_bound()  # :436,444
                self.elbo_vec[it] = lb
                if self.verbose:
                    print("iter: %d, ELBO: %.2f" % (it, lb))
                    sys.stdout.flush()
            else:
                self.iter_time[it] = time.time() - start_time
        import torch
        torch.cuda.synchronize(eng.device)
        self._host_stale = True
        self._pull_globals()
        self.metaobs_fun = None          # :485 (picklable)

    # ------------------------------------------------------------------ adaptive window machinery
    def _resize_locals(self, metaobs_sz):
        """hmmsgd_metaobs.py:360-367 / :382-388 (the random var_x keeps the RNG stream aligned)."""
        self.var_x = np.random.rand(metaobs_sz, self.K)
        self.var_x /= np.sum(self.var_x, axis=1)[:, np.newaxis]
        self.lalpha = np.empty((metaobs_sz, self.K))
        self.lbeta = np.empty((metaobs_sz, self.K))
        self.lliks = np.empty((metaobs_sz, self.K))

    def _window_marginals(self, inds, halflength):
        """Marginals (n, 2*halflength+1, K) of the windows centred at `inds`: ONE batched E-step
        instead of the reference's per-index get_local_messages calls (:663-700)."""
        eng = self._ensure_engine()
        inds = np.asarray(inds, dtype=np.int64)
        self._check_windows(inds - halflength, 2 * halflength + 1)
        vx, _ = eng.estep(inds - halflength, 2 * halflength + 1, flags=0)
        return vx.double().cpu().numpy()

    def get_local_messages(self, ind, halflength):
        """hmmsgd_metaobs.py:663-700: variational distribution over the window centred at ind, with
        the initial-state parameter self.var_init as it currently stands."""
        eng = self._ensure_engine()
        eng.set_var_init(self.var_init)
        try:
            return self._window_marginals([ind], halflength)[0]
        finally:
            eng.set_var_init(None if not self._explicit_init else self.var_init)

    def get_marginal(self, var_over_x, index):
        """hmmsgd_metaobs.py:702-708."""
        return np.squeeze(var_over_x[index, :])

    def select_L(self, numIndices=1, epsilon=1e-5, minHalfL=1, avgResidual=False, Lincrement=1, Lcutoff=1000):
        """hmmsgd_metaobs.py:521-569: grow the half-length until the centre marginal moves less than
        epsilon in L1.  All still-growing indices share the same L in every round, so each round is one
        batched E-step over them."""
        indices = npr.choice(self.T - 2 * minHalfL - 1, size=numIndices) + minHalfL
        eng = self._ensure_engine()
        eng.set_var_init(self.var_init)
        try:
            n = len(indices)
            L = minHalfL
            q_old = self._window_marginals(indices, L)[:, L]
            finalL = np.full(n, minHalfL)
            active = np.ones(n, dtype=bool)
            q_diff = np.full(n, np.finfo(np.float64).max)
            count = np.zeros(n, dtype=int); run_av = np.zeros(n); run_old = np.zeros(n)
            while active.any():
                grow = active & ~((indices - L < 1 + Lincrement) | (indices + L + Lincrement + 1 > self.T) | (L > Lcutoff))
                if not avgResidual:
                    grow &= ~(q_diff < epsilon)
                else:
                    count[grow] += 1
                    with np.errstate(divide='ignore', invalid='ignore'):
                        conv = (count > 1) & ((run_av - run_old) / np.maximum(count - 1, 1) < epsilon)
                    grow &= ~conv
                finalL[active & ~grow] = L
                active = grow
                if not active.any():
                    break
                L += Lincrement
                ia = np.nonzero(active)[0]
                q_new = self._window_marginals(indices[ia], L)[:, L]
                d = np.sum(np.abs(q_new - q_old[ia]), axis=1)
                if not avgResidual:
                    q_diff[ia] = d
                else:
                    run_old[ia] = run_av[ia]
                    run_av[ia] += d
                q_old[ia] = q_new
            return int(finalL.max())
        finally:
            eng.set_var_init(None if not self._explicit_init else self.var_init)

    def buffer_budget(self, halfL, budget=400):
        """hmmsgd_metaobs.py:571-577."""
        return int(np.ceil(budget / (2 * halfL + 1)))

    def select_buffer(self, numIndices=1, epsilon=1e-5, halfL=10, avgResidual=False, Lincrement=1, Lcutoff=1000):
        """hmmsgd_metaobs.py:579-661: grow the buffer until the marginals at the two endpoints of the
        original meta-observation move less than epsilon.  (The reference's avgResidual branch reads
        an undefined var_new, :642-643; here it is computed like in the plain branch.)"""
        indices = npr.choice(self.T - 2 * halfL - 1, size=numIndices) + halfL
        eng = self._ensure_engine()
        eng.set_var_init(self.var_init)
        try:
            n = len(indices)
            bufL = halfL
            v = self._window_marginals(indices, bufL)
            ql, qr = v[:, bufL - halfL].copy(), v[:, bufL + halfL].copy()
            finalL = np.full(n, halfL)
            active = np.ones(n, dtype=bool)
            dl = np.full(n, np.finfo(np.float64).max); dr = dl.copy()
            count = np.zeros(n, dtype=int)
            avl = np.zeros(n); avr = np.zeros(n); oldl = np.zeros(n); oldr = np.zeros(n)
            while active.any():
                grow = active & ~((indices - bufL < 1 + Lincrement) | (indices + bufL + Lincrement + 1 > self.T) | (bufL > Lcutoff))
                if not avgResidual:
                    grow &= ~((dl < epsilon) & (dr < epsilon))
                else:
                    count[grow] += 1
                    cm = np.maximum(count - 1, 1)
                    grow &= ~((count > 1) & ((avl - oldl) / cm < epsilon) & ((avr - oldr) / cm < epsilon))
                finalL[active & ~grow] = bufL
                active = grow
                if not active.any():
                    break
                bufL += Lincrement
                ia = np.nonzero(active)[0]
                v = self._window_marginals(indices[ia], bufL)
                nl, nr = v[:, bufL - halfL], v[:, bufL + halfL]
                d_l, d_r = np.sum(np.abs(nl - ql[ia]), axis=1), np.sum(np.abs(nr - qr[ia]), axis=1)
                if not avgResidual:
                    dl[ia], dr[ia] = d_l, d_r
                else:
                    oldl[ia], oldr[ia] = avl[ia], avr[ia]
                    avl[ia] += d_l; avr[ia] += d_r
                ql[ia], qr[ia] = nl, nr
            return int(finalL.max())
        finally:
            eng.set_var_init(None if not self._explicit_init else self.var_init)

    def intermediate_pars_buffer(self, metaobs, bufferL, L):
        """hmmsgd_metaobs.py:932-1008: E-step on the buffered window, statistics from its inner 2L+1
        rows (engine: svihmm_estep_buffered with trim = bufferL - L)."""
        eng = self._ensure_engine()
        T = metaobs.i2 - metaobs.i1 + 1
        vx, stats = eng.estep([metaobs.i1], T, flags=self._flags(), trim=bufferL - L)
        self._last_stats_host = eng.unpack_stats(stats)
        self.var_x = vx[0].double().cpu().numpy()
        return self.intermediate_pars(metaobs)

    def local_update(self, metaobs=None):
        """hmmsgd_metaobs.py:487-519 for one meta-observation (or the full series)."""
        if metaobs is None:
            metaobs = MetaObs(0, self.T - 1)
        eng = self._ensure_engine()
        T = metaobs.i2 - metaobs.i1 + 1
        vx, stats = eng.estep([metaobs.i1], T, flags=self._flags(), keep_locals=True)
        self._stats = stats
        self._pull_globals()
        self.var_init = eng.get_globals()[1]        # the stationary vector that was used (:418)
        self._materialise_locals(eng, vx, 1, T)
        self._last_stats_host = eng.unpack_stats(stats)

    def forward_msgs(self, metaobs=None):
        self.local_update(metaobs)

    def backward_msgs(self, metaobs=None):
        self.local_update(metaobs)

    def intermediate_pars(self, metaobs=None):
        """hmmsgd_metaobs.py:857-928: (A_inter, emit_inter) of the last local_update."""
        s = self._last_stats_host
        if self.emission_kind == "categorical":         # :907-926: alpha posterior - 1 per window
            emit_inter = [np.asarray(self.prior_emit[k].alphav_0) + s["sx"][k] - 1. for k in range(self.K)]
            return s["A"], emit_inter
        emit_inter = [[s["sx"][k], s["n"][k], s["sxx"][k], s["n"][k]] for k in range(len(s["n"]))]
        return s["A"], emit_inter      # mixtures: one entry per component, row k*C + c

    def global_update(self, A_inter, emit_inter=None):
        """hmmsgd_metaobs.py:1010-1069.  Accepts the packed device statistics (engine path) or the
        reference's (A_inter, emit_inter) pair, which is packed and shipped to the device."""
        import torch
        eng = self._ensure_engine()
        Lh, S, T = self.metaobs_half, self.mb_sz, self.T
        bfact_A = (T - 2 * Lh - 1) / (2. * Lh * S)                       # :1033
        bfact_E = (T - 2 * Lh - 1) / ((2. * Lh + 1.) * S)                # :1048
        if isinstance(A_inter, torch.Tensor):
            stats = A_inter
        else:
            stats = torch.from_numpy(self._pack_inter(A_inter, emit_inter)).to(eng.device)
            if stats.numel() != eng.slen:
                raise RuntimeError("packed statistics have %d entries, the engine expects %d" % (
                    stats.numel(), eng.slen))
        if self._px is not None and isinstance(A_inter, torch.Tensor):
            self._px.global_update(stats, self.lrate, bfact_A, bfact_E)
            self._px.reduced_stats(stats)   # callers read the all-reduced statistics from `stats`
        else:
            eng.global_update(stats, self.lrate, bfact_A, bfact_E)
        self._host_stale = True

    def _pack_inter(self, A_inter, emit_inter):
        """The reference's (A_inter, emit_inter) pair -> packed statistics of include/svihmm.h
        [A | n | sx | sxx | q0 | logZ, Q4, B, 0].  A_inter already holds the per-window prior terms
        (quirk Q5), so the tail carries the number of windows they stand for only where the device
        update needs it (Categorical: alpha_0 - 1 per window, :925-926)."""
        K = self.K
        A = np.asarray(A_inter, dtype=float).ravel()
        if A.size != K * K:
            raise RuntimeError("A_inter must be (K, K)")
        tail = np.zeros(K + 4)
        if self.emission_kind == "categorical":
            # emit_inter[k] = sum over windows of (alphav_0 + counts - 1)  (:925-926)
            nwin = float(getattr(self, "_last_B_global", self.mb_sz))
            a0 = np.array([np.asarray(self.prior_emit[k].alphav_0, dtype=float) for k in range(K)])
            counts = np.array([np.asarray(e, dtype=float).ravel() for e in emit_inter]) - nwin * (a0 - 1.)
            tail[K + 2] = nwin
            return np.concatenate([A, counts.sum(1), counts.ravel(), tail])
        if self.emission_kind not in ("niw_full", "niw_diag"):
            raise RuntimeError("global_update(A_inter, emit_inter): unsupported emission kind %r" % (
                self.emission_kind,))
        KE = len(emit_inter)
        if KE != self._ensure_engine().KE:
            raise RuntimeError("emit_inter has %d entries, expected %d" % (KE, self._ensure_engine().KE))
        return np.concatenate([A, np.array([float(e[1]) for e in emit_inter]),
                               np.concatenate([np.asarray(e[0], dtype=float).ravel() for e in emit_inter]),
                               np.concatenate([np.asarray(e[2], dtype=float).ravel() for e in emit_inter]),
                               tail])

    def pred_logprob(self, metaobs=None):
        """hmmsgd_metaobs.py:1086-1119: mean over the masked rows of the meta-observation of
        logsumexp_k( log(var_x[t,k] + eps) + E[log p(x_t | k)] ) with x_t the unmasked observation."""
        metaobs = self.cur_mo if metaobs is None else metaobs
        self.local_update(metaobs=metaobs)
        m = np.asarray(self.mask[metaobs.i1:metaobs.i2 + 1], dtype=bool)
        if m.sum() == 0:
            return None
        lp = np.log(self.var_x[m] + eps) + self.lliks[m]
        return float(np.mean(np.logaddexp.reduce(lp, axis=1)))

    def pred_logprob_full(self):
        """hmmsgd_metaobs.py:1121-1145: the same over all masked rows of the series, with var_x from
        full_local_update (masked rows carry no evidence there)."""
        m = np.asarray(self.mask, dtype=bool)
        if m.sum() == 0:
            return None
        q = self.full_local_update()
        eng = self._ensure_engine()
        eng.estep([0], self.T, flags=0, want_var_x=False, keep_locals=True)     # lliks of the unmasked observations
        ll = eng.get_locals(1, self.T)["lliks"][0]
        lp = np.log(q[m] + eps) + ll[m]
        return float(np.mean(np.logaddexp.reduce(lp, axis=1)))

    def full_local_update(self):
        """hmmsgd_metaobs.py:1147-1205: posterior over the whole series with masked rows carrying
        no evidence (obs[mask] = nan -> ll = 0)."""
        eng = self._ensure_engine()
        vx, _ = eng.estep([0], self.T, flags=L.MASK_LL)
        return vx[0].double().cpu().numpy()
