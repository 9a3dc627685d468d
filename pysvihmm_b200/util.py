"""Host-side helpers mirroring the reference's util.py (parameter bookkeeping only; the weighted
statistics of util.NIW_suffstats, util.py:73-83, are computed on the GPU by the engine)."""
import numpy as np


def NIW_zero_nat_pars(G):
    """util.py:12-14."""
    p = len(G.mu_mf)
    return [np.zeros(p), 0., np.zeros((p, p)), 0.]


def NIW_mf_natural_pars(mu, sigma, kappa, nu):
    """util.py:28-37 (eta3 = sigma + kappa mu mu^T, as the reference code does)."""
    p = len(mu)
    return [kappa * mu, kappa, sigma + np.outer(mu, mu) * kappa, nu + 2 + p]


def NIW_mf_moment_pars(G, e1, e2, e3, e4):
    """util.py:40-60."""
    p = len(e1)
    mu = e1 / e2
    G.mu_mf, G.sigma_mf, G.kappa_mf, G.nu_mf = mu, e3 - np.outer(mu, mu) * e2, e2, e4 - 2 - p
    G.mu = G.mu_mf
    G.sigma = G.sigma_mf / (G.nu_mf - p - 1)


def NIW_nat2moment_pars(e1, e2, e3, e4):
    """util.py:17-26 -- NOTE the reference subtracts outer(mu, mu) / kappa here (its comment says the
    convention "may be wrong"), unlike NIW_mf_moment_pars which subtracts outer(mu, mu) * kappa; both
    are kept as the reference has them.  Returns [mu, sigma, kappa, nu]."""
    p = len(e1)
    mu = e1 / e2
    return [mu, e3 - np.outer(mu, mu) / e2, e2, e4 - 2 - p]


def KL_gaussian(mu0, sig0, mu1, sig1):
    """util.py:87-103: KL(N(mu0, sig0) || N(mu1, sig1)); RuntimeError on mismatched shapes."""
    mu0, mu1, sig0, sig1 = np.asarray(mu0), np.asarray(mu1), np.asarray(sig0), np.asarray(sig1)
    D = len(mu0)
    if len(mu1) != D or sig0.shape[0] != D or sig1.shape[0] != D:
        raise RuntimeError("Means and covariances my be the same dimension.")
    if sig0.shape[0] != sig0.shape[1] or sig1.shape[0] != sig1.shape[1]:
        raise RuntimeError("Covariance matrices must be square.")
    dx = mu1 - mu0
    quad = dx.dot(np.linalg.solve(sig1, dx))
    tr = np.trace(np.linalg.solve(sig1, sig0))
    return 0.5 * (tr + quad - D - np.linalg.slogdet(sig0)[1] + np.linalg.slogdet(sig1)[1])


def mvnrand(mean, cov, size=1):
    """util.py:148-161: `size` draws of N(mean, cov) from the legacy global numpy RNG (same draw order
    as the reference: one randn(size, D) block), squeezed."""
    mu = np.squeeze(mean)
    z = np.random.randn(size, mu.shape[0])
    return np.squeeze(mu + z.dot(np.linalg.cholesky(cov).T))


def match_state_seq(sts_true, sts_pred, K):
    """util.py:210-234: exhaustive search over the K! relabellings for the one with the smallest
    Hamming distance (first minimum in itertools.permutations order); perm[pred] ~ true."""
    import itertools
    sts_true = np.asarray(sts_true).astype(int)
    sts_pred = np.asarray(sts_pred).astype(int)
    best, best_hd = None, np.inf
    for cand in itertools.permutations(range(K)):
        hd = np.mean(np.asarray(cand)[sts_pred] != sts_true)
        if hd < best_hd:
            best, best_hd = cand, hd
    return np.array(best)


def dirichlet_natural_pars(alpha):
    return alpha - 1.


def dirichlet_moment_pars(eta):
    return eta + 1.


def make_mask(sts, miss=0., left=0):
    """util.py:163-191: mark a `miss` fraction of each state's observations (right of `left`)
    as missing, drawn with the legacy global numpy RNG like the reference."""
    sts_l = sts[left:]
    K = np.unique(sts_l).shape[0]
    mask = np.zeros(len(sts), dtype='bool')
    if miss > 0.:
        for k in range(K):
            obs_k = np.where(sts_l == k)[0]
            if obs_k.shape[0] < 10:
                continue
            nobs_k = np.ceil(miss * np.sum(sts == k))
            if obs_k.shape[0] < nobs_k:
                nobs_k = np.ceil(miss * obs_k.shape[0])
            inds = np.random.choice(obs_k, size=int(nobs_k), replace=False)
            mask[left + inds] = True
    return mask


def make_mask_prediction(sts, miss=0.):
    """util.py:194-206."""
    nobs = len(sts)
    mask = np.zeros(nobs, dtype='bool')
    if miss == 0.:
        return mask
    mask[-int(np.ceil(miss * nobs)):] = True
    return mask


def munkres_match(sts_true, sts_pred, K):
    """util.py:236-277 via scipy's Hungarian solver: permutation perm with perm[pred] = true."""
    from scipy.optimize import linear_sum_assignment
    cost = np.zeros((K, K))
    for i in range(K):
        for j in range(K):
            cost[i, j] = np.sum(np.logical_and(sts_true == i, sts_pred != j))
    rows, cols = linear_sum_assignment(cost)
    perm = np.zeros(K, dtype=int)
    perm[cols] = rows
    return perm
