"""Parameter holders with the attribute surface of the pybasicbayes emission objects that the
reference's HMM drivers touch (mu_mf, sigma_mf, kappa_mf, nu_mf, mu_0, ...;
pybasicbayes/distributions.py:179-212).  The mean-field arithmetic itself
(expected_log_likelihood :351-366, weighted statistics :240-263, conjugate update :265-276)
runs on the GPU inside the engine; expected_log_likelihood here dispatches to it."""
import numpy as np
from scipy import special


class Gaussian(object):
    """Full-covariance Gaussian with a Normal-inverse-Wishart mean-field posterior."""
    kind = "niw_full"

    def __init__(self, mu=None, sigma=None, mu_0=None, sigma_0=None, kappa_0=None, nu_0=None,
                 kappa_mf=None, nu_mf=None):
        self.mu, self.sigma = mu, sigma
        self.mu_0, self.sigma_0, self.kappa_0, self.nu_0 = mu_0, sigma_0, kappa_0, nu_0
        self.kappa_mf = kappa_mf if kappa_mf is not None else kappa_0
        self.nu_mf = nu_mf if nu_mf is not None else nu_0
        self.mu_mf, self.sigma_mf = mu, sigma
        if mu is None and sigma is None and all(v is not None for v in (mu_0, sigma_0, kappa_0, nu_0)):
            self.resample()

    @property
    def hypparams(self):
        return dict(mu_0=self.mu_0, sigma_0=self.sigma_0, kappa_0=self.kappa_0, nu_0=self.nu_0)

    def rvs(self, size=None):
        """distributions.py:123-126 (legacy global RNG, same draw order)."""
        size = 1 if size is None else size
        size = size + (self.mu.shape[0],) if isinstance(size, tuple) else (size, self.mu.shape[0])
        return self.mu + np.random.normal(size=size).dot(np.linalg.cholesky(self.sigma).T)

    def resample(self, data=[]):
        """Random initialisation from the NIW prior (distributions.py:293-297).  The sampler lives
        in the reference's absent pymattutil dependency: standard Bartlett/QR construction here,
        results equal in distribution only (parity unpinned)."""
        D = len(self.mu_0)
        n = int(self.nu_0) if self.nu_0 == np.round(self.nu_0) else None
        chol = np.linalg.cholesky(self.sigma_0)
        if n is not None:
            x = np.random.randn(n, D)
        else:
            x = np.diag(np.sqrt(np.atleast_1d(np.random.chisquare(self.nu_0 - np.arange(D)))))
            x[np.triu_indices_from(x, 1)] = np.random.randn(D * (D - 1) // 2)
        R = np.linalg.qr(x, 'r')
        Tm = np.linalg.solve(R.T, chol.T).T
        sigma = Tm.dot(Tm.T)
        mu = np.random.multivariate_normal(self.mu_0, sigma / self.kappa_0)
        self.mu_mf, self.sigma_mf = self.mu, self.sigma = mu, sigma
        return self

    def expected_log_likelihood(self, x):
        """distributions.py:351-359, evaluated by the CUDA emission kernel (no CPU path)."""
        from .engine import EStepEngine
        x = np.ascontiguousarray(np.reshape(np.asarray(x, dtype=np.float64), (-1, len(self.mu_mf))))
        eng = EStepEngine(1, x.shape[1], self.kind)
        try:
            eng.set_series(x)
            eng.set_globals(np.ones((1, 1)), eng.pack_emit(*self._mf_arrays()))
            eng.estep([0], x.shape[0], want_var_x=False, keep_locals=True)
            return eng.get_locals(1, x.shape[0])["lliks"][0, :, 0]
        finally:
            eng.close()

    def _mf_arrays(self):
        return (np.asarray(self.mu_mf)[None], np.asarray(self.sigma_mf)[None],
                np.asarray(self.kappa_mf)[None], np.asarray(self.nu_mf)[None])

    def _prior_arrays(self):
        return (np.asarray(self.mu_0)[None], np.asarray(self.sigma_0)[None],
                np.asarray(self.kappa_0)[None], np.asarray(self.nu_0)[None])

    def _loglmbdatilde(self):
        """distributions.py:361-366 (host copy used only by the ELBO diagnostic get_vlb)."""
        D = len(self.mu_0)
        chol = np.linalg.cholesky(self.sigma_mf)
        return special.digamma((self.nu_mf - np.arange(D)) / 2.).sum() + D * np.log(2) \
            - 2 * np.log(chol.diagonal()).sum()

    def get_vlb(self):
        """distributions.py:331-349 (ELBO diagnostic; the inverse-Wishart entropy / partition
        function come from the absent pymattutil and are restated from their textbook form:
        parity unpinned)."""
        D = len(self.mu_0)
        llt = self._loglmbdatilde()
        dmu = self.mu_mf - self.mu_0
        q_entropy = -0.5 * (llt + D * (np.log(self.kappa_mf / (2 * np.pi)) - 1)) \
            + _iw_entropy(self.sigma_mf, self.nu_mf)
        p_avgengy = 0.5 * (D * np.log(self.kappa_0 / (2 * np.pi)) + llt - D * self.kappa_0 / self.kappa_mf
                           - self.kappa_0 * self.nu_mf * np.dot(dmu, np.linalg.solve(self.sigma_mf, dmu))) \
            + _iw_logpartition(self.sigma_0, self.nu_0) + (self.nu_0 - D - 1) / 2 * llt \
            - 0.5 * self.nu_mf * np.linalg.solve(self.sigma_mf, self.sigma_0).trace()
        return p_avgengy + q_entropy


class DiagonalGaussian(Gaussian):
    """EXTENSION (BASELINE config 2): D independent one-dimensional NIW factors per state.  The
    reference's DiagonalGaussian (distributions.py:667-797) has no mean-field path; the update
    rules are Gaussian's (:351-366, util.py:28-60) applied per dimension with D = 1.
    sigma / sigma_0 / *_mf are length-D vectors of per-dimension scale parameters;
    kappa and nu may be scalars or length-D vectors."""
    kind = "niw_diag"

    def rvs(self, size=None):
        size = 1 if size is None else size
        size = size + (self.mu.shape[0],) if isinstance(size, tuple) else (size, self.mu.shape[0])
        return self.mu + np.random.normal(size=size) * np.sqrt(self.sigma)

    def resample(self, data=[]):
        D = len(self.mu_0)
        nu = np.broadcast_to(np.asarray(self.nu_0, dtype=float), (D,))
        sig = np.asarray(self.sigma_0) / np.random.chisquare(nu)
        mu = self.mu_0 + np.random.randn(D) * np.sqrt(sig / self.kappa_0)
        self.mu_mf, self.sigma_mf = self.mu, self.sigma = mu, sig
        return self

    def get_vlb(self):
        D = len(self.mu_0)
        ka = np.broadcast_to(np.asarray(self.kappa_mf, dtype=float), (D,))
        nu = np.broadcast_to(np.asarray(self.nu_mf, dtype=float), (D,))
        ka0 = np.broadcast_to(np.asarray(self.kappa_0, dtype=float), (D,))
        nu0 = np.broadcast_to(np.asarray(self.nu_0, dtype=float), (D,))
        tot = 0.
        for d in range(D):
            g = Gaussian(mu=self.mu_mf[d:d + 1], sigma=np.array([[self.sigma_mf[d]]]),
                         mu_0=np.asarray(self.mu_0)[d:d + 1], sigma_0=np.array([[self.sigma_0[d]]]),
                         kappa_0=ka0[d], nu_0=nu0[d], kappa_mf=ka[d], nu_mf=nu[d])
            tot += g.get_vlb()
        return tot


class Categorical(object):
    """Categorical over symbols 0..C-1 with a Dirichlet mean-field posterior
    (pybasicbayes/distributions.py:1273-1418): the attribute surface the HMM drivers touch
    (alphav_0, _alpha_mf, weights, num_parameters, expected_log_likelihood, meanfieldupdate, get_vlb).
    Data are symbol indices, not indicator vectors (:1276-1283)."""
    kind = "categorical"

    def __init__(self, weights=None, alpha_0=None, K=None, alphav_0=None, alpha_mf=None):
        self.K = K
        self.alphav_0 = None if alphav_0 is None else np.asarray(alphav_0, dtype=np.float64)
        if self.alphav_0 is None and alpha_0 is not None and K is not None:
            self.alphav_0 = np.repeat(float(alpha_0) / K, K)              # :1308-1311
        if self.alphav_0 is not None:
            self.K = len(self.alphav_0)
        self.alpha_0 = alpha_0
        self._alpha_mf = np.asarray(alpha_mf, dtype=np.float64) if alpha_mf is not None else (
            np.asarray(weights, dtype=np.float64) * self.K if weights is not None else None)   # :1298
        self.weights = weights
        if weights is None and self.alphav_0 is not None:
            self.resample()                                               # :1302-1303

    @property
    def alpha_mf(self):
        return self._alpha_mf

    def num_parameters(self):
        return self.K

    def resample(self, data=[]):
        """Initialise from the Dirichlet prior (:1335-1340, no data)."""
        self.weights = np.random.dirichlet(self.alphav_0)
        if self._alpha_mf is None:
            self._alpha_mf = self.weights * self.alphav_0.sum()
        return self

    def rvs(self, size=None):
        return np.random.choice(self.K, size=size, p=self.weights)

    def expected_log_likelihood(self, x=None):
        """distributions.py:1383-1386, evaluated by the CUDA emission kernel (no CPU path)."""
        from .engine import EStepEngine
        x = np.arange(self.K) if x is None else np.asarray(x)
        xs = np.ascontiguousarray(x.reshape(-1, 1).astype(np.float64))
        eng = EStepEngine(1, self.K, self.kind)
        try:
            eng.set_series(xs)
            eng.set_globals(np.ones((1, 1)), self._alpha_mf[None])
            eng.estep([0], xs.shape[0], want_var_x=False, keep_locals=True)
            return eng.get_locals(1, xs.shape[0])["lliks"][0, :, 0].reshape(x.shape)
        finally:
            eng.close()

    def get_vlb(self):
        """distributions.py:1372-1381 (host: ELBO diagnostic)."""
        logpitilde = special.digamma(self._alpha_mf) - special.digamma(self._alpha_mf.sum())
        q_entropy = -1 * ((logpitilde * (self._alpha_mf - 1)).sum()
                          + special.gammaln(self._alpha_mf.sum()) - special.gammaln(self._alpha_mf).sum())
        p_avgengy = special.gammaln(self.alphav_0.sum()) - special.gammaln(self.alphav_0).sum() \
            + ((self.alphav_0 - 1) * logpitilde).sum()
        return p_avgengy + q_entropy


class MixtureDistribution(object):
    """EXTENSION (BASELINE config 5): a mixture of NIW Gaussians used as the emission of one HMM
    state, with the attribute surface of pybasicbayes.models.MixtureDistribution (models.py:256-300:
    `components`, `weights` = a Categorical).  The reference class has Gibbs/EM methods only; the
    mean-field E-step / statistics / natural-gradient step run in the engine (svihmm_create_mix)."""

    def __init__(self, components, weights=None, alpha_0=None):
        self.components = list(components)
        C = len(self.components)
        self.weights = weights if weights is not None else Categorical(
            weights=np.ones(C) / C, alphav_0=np.ones(C) * (1. if alpha_0 is None else float(alpha_0) / C),
            alpha_mf=np.ones(C))
        self.kind = self.components[0].kind

    def get_vlb(self):
        return self.weights.get_vlb() + sum(c.get_vlb() for c in self.components)


def _iw_logpartition(sigma, nu):
    D = sigma.shape[0]
    chol = np.linalg.cholesky(sigma)
    return -1 * (nu * np.log(chol.diagonal()).sum()
                 - (nu * D / 2 * np.log(2) + D * (D - 1) / 4 * np.log(np.pi)
                    + special.gammaln((nu - np.arange(D)) / 2).sum()))


def _iw_entropy(sigma, nu):
    D = sigma.shape[0]
    chol = np.linalg.cholesky(sigma)
    Elogdet = special.digamma((nu - np.arange(D)) / 2).sum() + D * np.log(2) - 2 * np.log(chol.diagonal()).sum()
    return _iw_logpartition(sigma, nu) - (nu - D - 1) / 2 * Elogdet + nu * D / 2
