"""hmmbatchsgd.VBHMM over the CUDA engine: batch natural-gradient VB on the full sequence
(reference hmmbatchsgd.py).  Same constructor and infer() control flow; every iteration runs the
full-sequence E-step (hmmbase.py:201-229, with the masked rows NaN-ed out as hmmbatchsgd.py:148-149
does) and the natural-gradient blend of hmmbatchsgd.py:202-259 on the GPU."""
import sys
import time

import numpy as np

from . import _lib as L
from .hmmbase import VariationalHMMBase

tau0 = 1.
kappa0 = 0.7


class VBHMM(VariationalHMMBase):
    """Batch stochastic-gradient variational inference for hidden Markov models (hmmbatchsgd.py:25)."""

    @staticmethod
    def make_param_dict(prior_init, prior_tran, prior_emit, tau=tau0, kappa=kappa0, mask=None):
        """hmmbatchsgd.py:38-45."""
        return {'prior_init': prior_init, 'prior_tran': prior_tran, 'prior_emit': prior_emit,
                'mask': mask, 'tau': tau, 'kappa': kappa}

    def __init__(self, obs, prior_init, prior_tran, prior_emit, tau=tau0, kappa=kappa0, mask=None,
                 init_init=None, init_tran=None, epsilon=1e-8, maxit=100, verbose=False, sts=None,
                 obs_dtype="f64", device=None):
        """hmmbatchsgd.py:47-141."""
        super(VBHMM, self).__init__(obs, prior_init, prior_tran, prior_emit, mask=mask,
                                    init_init=init_init, init_tran=init_tran, verbose=verbose, sts=sts,
                                    obs_dtype=obs_dtype, device=device)
        self.batch = self.obs
        self.elbo = -np.inf
        self.tau, self.kappa = tau, kappa
        self.lrate = tau ** (-kappa)
        self.epsilon, self.maxit = epsilon, maxit
        self.batchfactor = 1.
        self.var_x = np.ones((self.T, self.K)) / self.K                  # :126-127
        self.lalpha = np.empty((self.T, self.K))
        self.lbeta = np.empty((self.T, self.K))
        self.lliks = np.empty((self.T, self.K))
        self.mod_init = np.zeros(self.K)
        self.mod_tran = np.zeros((self.K, self.K))

    def _local_flags(self):
        # masked rows carry no evidence (:148-149) and the transition statistic is prior + sum (:219-225)
        return L.MASK_LL | L.ADD_PRIOR

    def infer(self):
        """hmmbatchsgd.py:143-200."""
        maxit = self.maxit
        self.elbo_vec = np.inf * np.ones(maxit)
        self.pred_logprob_mean = np.nan * np.ones(maxit)
        self.pred_logprob_std = np.nan * np.ones(maxit)
        self.iter_time = np.nan * np.ones(maxit)
        for it in range(maxit):
            start_time = time.time()
            self.lrate = (it + self.tau) ** (-self.kappa)                 # :165
            self.local_update()
            self.global_update()
            self.iter_time[it] = time.time() - start_time
            lb = self.lower_bound()
            if self.verbose:
                print("iter: %d, ELBO: %.2f" % (it, lb))
                sys.stdout.flush()
            self.elbo = lb
            self.elbo_vec[it] = lb
        lbidx = np.where(np.logical_not(np.isinf(self.elbo_vec)))[0]
        self.elbo_vec = self.elbo_vec[lbidx]
        self.pred_logprob_mean = self.pred_logprob_mean[lbidx]
        self.pred_logprob_std = self.pred_logprob_std[lbidx]
        self.iter_time = self.iter_time[lbidx]
        self._pull_globals()
        if self.sts is not None:
            self.hamming, self.perm = self.hamming_dist(self.var_x, self.sts)

    def global_update(self, batch=None):
        """hmmbatchsgd.py:202-259 from the statistics of the last local_update (device)."""
        eng = self._ensure_engine()
        eng.batchsgd_update(self._stats, self.lrate)
        self._host_stale = True
        self._pull_globals()
