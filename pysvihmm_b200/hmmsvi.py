"""hmmsvi.SVIHMM surface (reference hmmsvi.py:88-204) over the CUDA engine.

The reference class is a non-functional skeleton (wrong super().__init__ argument order
hmmsvi.py:62-63, np.squash :139, self.N undefined :193), so there is no runnable behaviour to
match; what is kept is its method surface -- infer(mb_gen, maxit), local_update(batch),
global_update(batch), update_lrate(it), allobs_batch(), generate_obs(T) -- with the semantics its
code states: a minibatch is an iterable of observation indices forming one contiguous window,
the local step is the E-step on it, the global step blends natural parameters with step lrate
(hmmsvi.py:115-178)."""
import numpy as np

from . import _lib as L
from .hmmbase import VariationalHMMBase


class SVIHMM(VariationalHMMBase):
    def __init__(self, prior_init, prior_tran, prior_emit, obs, tau=1., kappa=0.7, obs_dtype="f64",
                 device=None):
        """hmmsvi.py:42-86 (argument order of the reference kept: priors first, obs last)."""
        super(SVIHMM, self).__init__(obs, prior_init, prior_tran, prior_emit, obs_dtype=obs_dtype,
                                     device=device)
        self.batch = None
        self.elbo = -np.inf
        self.tau, self.kappa = tau, kappa
        self.lrate = 0.
        self.batchfactor = 1.
        self.var_init = np.array(prior_init, dtype=float).copy()
        self.var_tran = np.array(prior_tran, dtype=float).copy()
        self.var_x = None
        self._it = 0

    def allobs_batch(self):
        """hmmsvi.py:88-91."""
        yield range(self.T)

    def update_lrate(self, it):
        """hmmsvi.py:93-94 is a stub; the Robbins-Monro schedule of hmmsgd_metaobs.py:351 is used."""
        self.lrate = (it + self.tau) ** (-self.kappa)

    def infer(self, mb_gen, maxit=100):
        """hmmsvi.py:96-112."""
        for it in range(maxit):
            batches = mb_gen() if callable(mb_gen) else mb_gen
            for batch in batches:
                self.update_lrate(self._it)
                self.local_update(batch)
                self.global_update(batch)
                self._it += 1
        self._pull_globals()

    def local_update(self, batch=None):
        """E-step on the contiguous index window `batch` (hmmsvi.py:104-109)."""
        idx = np.arange(self.T) if batch is None else np.asarray(list(batch))
        if np.any(np.diff(idx) != 1):
            raise RuntimeError("a minibatch must be a contiguous index window")
        self.batch = (int(idx[0]), int(idx[-1]))
        eng = self._ensure_engine()
        vx, stats = eng.estep([self.batch[0]], len(idx), flags=L.ADD_PRIOR)
        self._stats = stats
        self.var_x = vx[0].double().cpu().numpy()

    def global_update(self, batch=None):
        """hmmsvi.py:115-178: eta <- (1-lrate) eta + lrate (prior + batchfactor * stats)."""
        eng = self._ensure_engine()
        eng.global_update(self._stats, self.lrate, self.batchfactor, self.batchfactor)
        self._host_stale = True

    def generate_obs(self, T):
        """hmmsvi.py:180-204 (with self.N read as K): sample states/observations from the prior
        mean transition matrix and the current emission parameters."""
        self._pull_globals()
        A = self.prior_tran / self.prior_tran.sum(1)[:, None]
        p0 = self.prior_init / self.prior_init.sum()
        st = np.random.choice(self.K, p=p0)
        sts, obs = [st], [self.var_emit[st].rvs()[0]]
        for _ in range(1, T):
            st = np.random.choice(self.K, p=A[st])
            sts.append(st)
            obs.append(self.var_emit[st].rvs()[0])
        return np.array(sts), np.array(obs)
