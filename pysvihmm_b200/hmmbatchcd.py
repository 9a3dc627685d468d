"""hmmbatchcd.VBHMM over the CUDA engine: batch coordinate-ascent VB on the full sequence
(reference hmmbatchcd.py; BASELINE config 1).  Same constructor and infer() control flow; the
E-step (hmmbase.py:201-229) and the conjugate global update (hmmbatchcd.py:172-189) run on the GPU."""
import sys
import time

import numpy as np

from .hmmbase import VariationalHMMBase

eps = 1e-9


class VBHMM(VariationalHMMBase):
    """Batch coordinate-descent variational inference for hidden Markov models (hmmbatchcd.py:22)."""

    def __init__(self, obs, prior_init, prior_tran, prior_emit, mask=None, init_init=None,
                 init_tran=None, epsilon=1e-8, maxit=100, verbose=False, sts=None,
                 obs_dtype="f64", device=None):
        """hmmbatchcd.py:44-108."""
        super(VBHMM, self).__init__(obs, prior_init, prior_tran, prior_emit, mask=mask,
                                    init_init=init_init, init_tran=init_tran, verbose=verbose, sts=sts,
                                    obs_dtype=obs_dtype, device=device)
        self.epsilon = epsilon
        self.maxit = maxit
        self.var_x = np.random.rand(self.T, self.K)              # :93-94
        self.var_x /= np.sum(self.var_x, axis=1)[:, np.newaxis]
        self.lalpha = np.empty((self.T, self.K))
        self.lbeta = np.empty((self.T, self.K))
        self.lliks = np.empty((self.T, self.K))
        self.mod_init = np.zeros(self.K)
        self.mod_tran = np.zeros((self.K, self.K))

    def infer(self):
        """hmmbatchcd.py:114-170."""
        epsilon, maxit = self.epsilon, self.maxit
        self.elbo_vec = np.inf * np.ones(maxit)
        self.pred_logprob_mean = np.nan * np.ones(maxit)
        self.pred_logprob_std = np.nan * np.ones(maxit)
        self.iter_time = np.nan * np.ones(maxit)
        for it in range(maxit):
            start_time = time.time()
            self.local_update()
            self.global_update()
            self.iter_time[it] = time.time() - start_time
            lb = self.lower_bound()
            if self.verbose:
                print("iter: %d, ELBO: %.2f" % (it, lb))
                sys.stdout.flush()
            if np.allclose(lb, self.elbo, atol=epsilon):            # :149
                break
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

    def global_update(self):
        """hmmbatchcd.py:172-189 from the statistics of the last local_update (device)."""
        eng = self._ensure_engine()
        eng.batch_update(self._stats)
        self._host_stale = True
        self._pull_globals()
