"""VariationalHMMBase: the reference's plugin surface (hmmbase.py:34-411) over the CUDA engine.

Same constructor, attributes and method names as the reference class; the E-step methods
(local_update, forward_msgs, backward_msgs) dispatch to libsvihmm.so instead of looping over t
in numpy.  State contract kept: obs, mask, K, T, D, prior_*, var_init, var_tran, var_emit,
mod_init, mod_tran, lliks, lalpha, lbeta, var_x, elbo.
"""
import abc
from copy import deepcopy

import numpy as np
from scipy.special import digamma, gammaln

from . import _lib as L
from . import util

eps = 1e-9   # hmmbase.py:30


class VariationalHMMBase(object, metaclass=abc.ABCMeta):
    """Abstract base class for finite variational HMMs (hmmbase.py:34)."""

    @abc.abstractmethod
    def global_update(self):
        pass

    @abc.abstractmethod
    def infer(self):
        pass

    @staticmethod
    def make_param_dict(prior_init, prior_tran, prior_emit, mask=None):
        """hmmbase.py:52-58."""
        return {'prior_init': prior_init, 'prior_tran': prior_tran,
                'prior_emit': prior_emit, 'mask': mask}

    def set_mask(self, mask):
        """hmmbase.py:60-65."""
        if mask is None:
            self.mask = np.zeros(self.obs.shape[0], dtype='bool')
        else:
            self.mask = np.asarray(mask).astype('bool')
        self._series_dirty = True

    def __init__(self, obs, prior_init, prior_tran, prior_emit, mask=None, init_init=None,
                 init_tran=None, verbose=False, sts=None, obs_dtype="f64", device=None):
        """hmmbase.py:67-136.  Extra engine kwargs: obs_dtype ('f64' keeps the reference's
        float64 series in HBM, 'f32' halves the traffic), device (CUDA index)."""
        self.verbose = verbose
        self.sts = sts
        self.prior_init = deepcopy(np.asarray(prior_init)).astype('float64')
        self.prior_tran = deepcopy(np.asarray(prior_tran)).astype('float64')
        self.prior_emit = deepcopy(prior_emit)
        if init_init is None:
            self.var_init = self.prior_init / np.sum(self.prior_init)
        else:
            self.var_init = np.array(init_init, dtype='float64')
        if init_tran is None:
            self.var_tran = self.prior_tran / np.sum(self.prior_tran, axis=1)[:, np.newaxis]
        else:
            self.var_tran = np.array(init_tran, dtype='float64')
        self.var_emit = deepcopy(prior_emit)
        self.obs = obs
        self.K = self.prior_tran.shape[0]
        if obs.ndim == 1:
            self.T, self.D = obs.shape[0], 1
        elif obs.ndim == 2:
            self.T, self.D = obs.shape
        else:
            raise RuntimeError("obs must have 1 or 2 dimensions")
        self.set_mask(mask)
        self.elbo = -np.inf
        self.obs_dtype = obs_dtype
        self._device = device
        self._engine = None
        self._series_dirty = True
        self._globals_dirty = True      # host copy newer than the device copy
        self._host_stale = False        # device copy newer than the host copy
        self._explicit_init = True      # var_init is an explicit Dirichlet parameter (batch drivers)

    def set_data(self, obs, mask=None):
        """hmmbase.py:138-143."""
        self.obs = obs
        self.set_mask(mask)

    # ------------------------------------------------------------------ engine plumbing
    @property
    def emission_kind(self):
        return getattr(self.var_emit[0], "kind", "niw_full")

    def _ensure_engine(self):
        from .engine import EStepEngine
        cat = self.emission_kind == "categorical"
        comps = getattr(self.var_emit[0], "components", None)
        if comps is not None:
            return self._ensure_engine_mix()
        if self._engine is None and cat:
            C = int(self.prior_emit[0].num_parameters())
            self._engine = EStepEngine(self.K, C, "categorical", device=self._device)
            self._engine.set_prior(self.prior_tran, np.array([np.asarray(g.alphav_0, dtype=float)
                                                              for g in self.prior_emit]), self.prior_init)
            self._series_dirty = True
            self._globals_dirty = True
        if self._engine is None:
            self._engine = EStepEngine(self.K, self.D, self.emission_kind, device=self._device)
            pe = self.prior_emit
            self._engine.set_prior(self.prior_tran, self._engine.pack_emit(
                np.array([np.asarray(g.mu_0, dtype=float) for g in pe]),
                np.array([np.asarray(g.sigma_0, dtype=float) for g in pe]),
                np.array([np.broadcast_to(np.asarray(g.kappa_0, dtype=float), self._kn_shape()) for g in pe]),
                np.array([np.broadcast_to(np.asarray(g.nu_0, dtype=float), self._kn_shape()) for g in pe])),
                self.prior_init)
            self._series_dirty = True
            self._globals_dirty = True
        if self._series_dirty:
            obs = np.asarray(self.obs, dtype=np.float64).reshape(self.T, -1)
            self._engine.set_series(obs, self.mask, dtype=self.obs_dtype)
            self._series_dirty = False
        if self._globals_dirty and cat:
            em = np.array([np.asarray(g._alpha_mf, dtype=float) for g in self.var_emit])
            self._engine.set_globals(self.var_tran, em, self.var_init if self._explicit_init else None)
            self._globals_dirty = False
        if self._globals_dirty:
            ve = self.var_emit
            em = self._engine.pack_emit(
                np.array([np.asarray(g.mu_mf, dtype=float) for g in ve]),
                np.array([np.asarray(g.sigma_mf, dtype=float) for g in ve]),
                np.array([np.broadcast_to(np.asarray(g.kappa_mf, dtype=float), self._kn_shape()) for g in ve]),
                np.array([np.broadcast_to(np.asarray(g.nu_mf, dtype=float), self._kn_shape()) for g in ve]))
            self._engine.set_globals(self.var_tran, em, self.var_init if self._explicit_init else None)
            self._globals_dirty = False
        return self._engine

    def _flat(self, emit):
        return [g for m in emit for g in m.components]

    def _ensure_engine_mix(self):
        """Mixture emissions: K*C component rows + Dirichlet weights (svihmm_create_mix)."""
        from .engine import EStepEngine
        C = len(self.var_emit[0].components)
        kshape = () if self.emission_kind == "niw_full" else (self.D,)
        bc = lambda v: np.broadcast_to(np.asarray(v, dtype=float), kshape)
        if self._engine is None:
            self._engine = EStepEngine(self.K, self.D, self.emission_kind, device=self._device, components=C)
            pe = self._flat(self.prior_emit)
            self._engine.set_prior(self.prior_tran, self._engine.pack_emit(
                np.array([np.asarray(g.mu_0, dtype=float) for g in pe]),
                np.array([np.asarray(g.sigma_0, dtype=float) for g in pe]),
                np.array([bc(g.kappa_0) for g in pe]), np.array([bc(g.nu_0) for g in pe])), self.prior_init)
            self._series_dirty = True
            self._globals_dirty = True
        if self._series_dirty:
            obs = np.asarray(self.obs, dtype=np.float64).reshape(self.T, -1)
            self._engine.set_series(obs, self.mask, dtype=self.obs_dtype)
            self._series_dirty = False
        if self._globals_dirty:
            ve = self._flat(self.var_emit)
            self._engine.set_mix_weights(np.array([m.weights._alpha_mf for m in self.var_emit]),
                                         np.array([m.weights.alphav_0 for m in self.prior_emit]))
            self._engine.set_globals(self.var_tran, self._engine.pack_emit(
                np.array([np.asarray(g.mu_mf, dtype=float) for g in ve]),
                np.array([np.asarray(g.sigma_mf, dtype=float) for g in ve]),
                np.array([bc(g.kappa_mf) for g in ve]), np.array([bc(g.nu_mf) for g in ve])),
                self.var_init if self._explicit_init else None)
            self._globals_dirty = False
        return self._engine

    def _kn_shape(self):
        return () if self.emission_kind == "niw_full" else (self.D,)

    def _pull_globals(self):
        """Device master copy -> host attributes (var_tran, var_init, var_emit[k].*_mf, mu, sigma)."""
        if self._engine is None or not self._host_stale:
            return
        vt, vi, em = self._engine.get_globals()
        self.var_tran, self.var_init = vt, vi
        e = self._engine.unpack_emit(em)
        mixed = getattr(self.var_emit[0], "components", None) is not None
        if mixed:
            om = self._engine.get_mix_weights()
            for k, M in enumerate(self.var_emit):
                M.weights._alpha_mf = om[k]
                M.weights.weights = om[k] / om[k].sum()
        if self.emission_kind == "categorical":
            for k, G in enumerate(self.var_emit):                 # hmmsgd_metaobs.py:1083-1084
                G._alpha_mf = e["alpha"][k]
                G.weights = G._alpha_mf / G._alpha_mf.sum()
            self._host_stale = False
            return
        full = self.emission_kind == "niw_full"
        for k, G in enumerate(self._flat(self.var_emit) if mixed else self.var_emit):
            G.mu_mf, G.sigma_mf = e["mu"][k], e["sigma"][k]
            G.kappa_mf = float(e["kappa"][k]) if full else e["kappa"][k]
            G.nu_mf = float(e["nu"][k]) if full else e["nu"][k]
            G.mu = G.mu_mf                                   # util.py:59-60
            G.sigma = G.sigma_mf / (G.nu_mf - (self.D if full else 1) - 1)
        self._host_stale = False

    # ------------------------------------------------------------------ E-step (hmmbase.py:201-229)
    def local_update(self, obs=None, mask=None):
        """Full-sequence E-step on the GPU: lliks, scaled forward/backward, var_x.
        Fills var_x (T,K), lliks, lalpha, lbeta (reconstructed log-domain tables), mod_init,
        mod_tran like hmmbase.py:201-229; keeps the packed statistics for global_update."""
        if obs is not None or mask is not None:
            self.set_data(self.obs if obs is None else obs, self.mask if mask is None else mask)
        eng = self._ensure_engine()
        vx, stats = eng.estep([0], self.T, flags=self._local_flags(), keep_locals=True)
        self._stats = stats
        self._materialise_locals(eng, vx, 1, self.T)

    def _local_flags(self):
        return 0

    def _materialise_locals(self, eng, vx, B, T):
        loc = eng.get_locals(B, T)
        self.var_x = vx[B - 1].double().cpu().numpy()
        self.lliks = loc["lliks"][B - 1]
        self.lalpha, self.lbeta = eng.log_tables(loc, B - 1)
        self.mod_init = digamma(self.var_init + eps) - digamma(np.sum(self.var_init) + eps)
        tran_sum = np.sum(self.var_tran, axis=1)
        self.mod_tran = digamma(self.var_tran + eps) - digamma(tran_sum[:, None] + eps)
        self._logZ, self._lb_q4 = loc["logZ"], loc["lb_q4"]

    def forward_msgs(self, obs=None, mask=None):
        """hmmbase.py:266-295.  The engine fuses forward/backward; calling either runs the
        E-step and fills lalpha."""
        self.local_update(obs, mask)

    def backward_msgs(self, obs=None, mask=None):
        """hmmbase.py:297-320 (see forward_msgs)."""
        self.local_update(obs, mask)

    def full_local_update(self):
        """hmmbase.py:342-344."""
        self.local_update()
        return self.var_x

    # ------------------------------------------------------------------ ELBO (hmmbase.py:145-199)
    def _dirichlet_bound(self, p, q):
        q_dg = digamma(q + eps)
        dg_sum = digamma(np.sum(q, axis=-1) + eps)[..., None]
        energy = gammaln(np.sum(p, axis=-1) + eps) - np.sum(gammaln(p + eps), axis=-1) \
            + np.sum((p - 1.) * (q_dg - dg_sum), axis=-1)
        entropy = -(gammaln(np.sum(q, axis=-1) + eps) - np.sum(gammaln(q + eps), axis=-1)
                    + np.sum((q - 1.) * (q_dg - dg_sum), axis=-1))
        return np.sum(energy) + np.sum(entropy)

    def lower_bound(self):
        """hmmbase.py:145-199; the data term lZ = sum_t logsumexp_k lalpha[t] comes from the
        engine (statistics tail)."""
        return self._ensure_engine().global_bound(include_init=True) + float(np.sum(self._lb_q4))

    def _lower_bound_host(self):
        """The same from the host copies of the parameters (cross-check of svihmm_global_bound)."""
        self._pull_globals()
        elbo = self._dirichlet_bound(self.prior_init, self.var_init)
        elbo += self._dirichlet_bound(self.prior_tran, self.var_tran)
        for k in range(self.K):
            elbo += self.var_emit[k].get_vlb()
        return elbo + float(np.sum(self._lb_q4))

    # ------------------------------------------------------------------ FFBS (hmmbase.py:231-264, hmm_fast.pyx)
    def FFBS(self, var_init, seed=None):
        """Forward-filter backward-sampling of one state path (hmmbase.py:231-264; the sampling loop
        follows the native version hmm_fast.pyx:103-122, the numpy one never refreshes p, :260-262)."""
        return self.ffbs_fast(var_init, seed=seed)[0]

    def ffbs_fast(self, var_init, lalpha_init=None, seed=None):
        """hmm_fast.FFBS (bound at hmmbase.py:410-411): returns (z, lalpha).  lalpha_init is accepted for
        signature compatibility; the filter is recomputed on the device (it costs one forward pass)."""
        eng = self._ensure_engine()
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1))      # tied to the legacy global RNG like rand()
        z = eng.ffbs(np.asarray(var_init, dtype=np.float64), nsamples=1, seed=seed)[0].astype(np.int64)
        loc = eng.get_locals(1, self.T)
        with np.errstate(divide='ignore'):
            cum = np.cumsum(np.log(loc["cs"][0].astype(np.float64)) + loc["mx"][0])
            lalpha = np.log(loc["alpha"][0].astype(np.float64)) + cum[:, None]
        return z, lalpha

    # ------------------------------------------------------------------ metrics
    def hamming_dist(self, full_var_x, true_sts):
        """hmmbase.py:346-362."""
        import scipy.spatial.distance as dist
        state_sq = np.argmax(full_var_x, axis=1).astype(int)
        best_match = util.munkres_match(true_sts, state_sq, self.K)
        return dist.hamming(true_sts, best_match[state_sq]), best_match

    def A_dist(self, A_true, perm):
        """hmmbase.py:392-406."""
        self._pull_globals()
        A = self.var_tran / np.sum(self.var_tran, axis=1)[:, np.newaxis]
        return np.linalg.norm(A_true[np.ix_(perm, perm)] - A)
