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
from .sharding import allreduce_stats, dist_or_none as _dist, shard_starts

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
                 track_elbo=True, pairwise_mode="ref_outer"):
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
        if adagrad:
            raise NotImplementedError("adagrad variant (hmmsgd_metaobs.py:1036-1040) is not on the engine yet")
        self.adagrad = adagrad
        self.maxit = maxit
        if growBuffer or bufferBudget:
            raise NotImplementedError("growBuffer/bufferBudget (hmmsgd_metaobs.py:579-661,932-1008) "
                                      "are 'next' rows of the scope table, not built yet")
        self.growBuffer = growBuffer
        self.bufferBudget = bufferBudget
        if metaobs_half < 1:
            raise RuntimeError("metaobs (%d) must be >= 1." % (metaobs_half,))
        self.metaobs_half = metaobs_half
        self.mb_sz = mb_sz
        self.cur_mo = None
        self.batchfactor = 1.
        self.track_elbo = track_elbo
        if pairwise_mode not in ("ref_outer", "exact_xi"):
            raise RuntimeError("pairwise_mode must be 'ref_outer' or 'exact_xi'")
        self.pairwise_mode = pairwise_mode
        metaobs_sz = 2 * metaobs_half + 1
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
        """hmmsgd_metaobs.py:273-296."""
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

    def minibatch_estep(self, minibatch, want_var_x=True):
        """Batched replacement of the loop hmmsgd_metaobs.py:405-436 for this rank's share of the
        minibatch; returns the packed, all-reduced statistics tensor (device)."""
        eng = self._ensure_engine()
        starts = np.array([m.i1 for m in minibatch], dtype=np.int64)
        T = int(minibatch[0].i2 - minibatch[0].i1 + 1)
        dist = _dist()
        if dist is not None:
            starts = shard_starts(starts, dist.get_rank(), dist.get_world_size())
        vx, stats = eng.estep(starts, T, flags=self._flags(), want_var_x=want_var_x)
        allreduce_stats(stats, dist)        # one sum all-reduce of the packed statistics per step
        self._var_x_batch = vx
        self._last_B, self._last_T = len(starts), T
        return stats

    def infer(self, adaptive=False, perIter=10, epsilon=1e-6, minHalfL=1, avgResidual=False,
              Lincrement=1, Lcutoff=1000):
        """hmmsgd_metaobs.py:298-485."""
        if adaptive or self.metaobs_half is None:
            raise NotImplementedError("select_L (hmmsgd_metaobs.py:521-569) is a 'next' row, not built yet")
        np.random.seed(self.seed)
        maxit = self.maxit
        if self.metaobs_fun is None:
            self.set_metaobs_fun()
        self.elbo_vec = np.inf * np.ones(maxit)
        self.iter_time = np.inf * np.ones(maxit)
        mb_sz, Lh = self.mb_sz, self.metaobs_half
        eng = self._ensure_engine()
        for it in range(maxit):
            start_time = time.time()
            self.lrate = (it + self.tau) ** (-self.kappa)                 # :351
            minibatch = self.metaobs_fun(self.T, Lh, mb_sz)               # :396
            self.cur_mo = minibatch[-1]
            stats = self.minibatch_estep(minibatch, want_var_x=False)     # :405-436
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
            K, D = self.K, self.D
            parts = [np.asarray(A_inter, dtype=float).ravel(),
                     np.array([e[1] for e in emit_inter], dtype=float),
                     np.concatenate([np.asarray(e[0], dtype=float).ravel() for e in emit_inter]),
                     np.concatenate([np.asarray(e[2], dtype=float).ravel() for e in emit_inter]),
                     np.zeros(K + 4)]
            stats = torch.from_numpy(np.concatenate(parts)).to(eng.device)
        eng.global_update(stats, self.lrate, bfact_A, bfact_E)
        self._host_stale = True

    def full_local_update(self):
        """hmmsgd_metaobs.py:1147-1205: posterior over the whole series with masked rows carrying
        no evidence (obs[mask] = nan -> ll = 0)."""
        eng = self._ensure_engine()
        vx, _ = eng.estep([0], self.T, flags=L.MASK_LL)
        return vx[0].double().cpu().numpy()
