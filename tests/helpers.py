"""Shared test helpers: golden fixture loading and packing of emission params."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def emit_list(mu, sigma, kappa, nu):
    return [dict(mu=mu[k], sigma=sigma[k], kappa=float(kappa[k]), nu=float(nu[k]))
            for k in range(len(mu))]


def golden_prior_emit(g, K):
    return [dict(mu=g["prior_mu"], sigma=g["prior_sigma"], kappa=float(g["prior_kappa"]),
                 nu=float(g["prior_nu"])) for _ in range(K)]


SVI_CASES = ["svi_k3_d2_l5", "svi_k5_d3_l20_mask", "svi_k16_d8_l50", "svi_k2_d2_l1"]


def frac_soft(q):
    """Fraction of timesteps whose posterior is not one-hot (vacuous-parity guard)."""
    return float(np.mean(np.max(q, axis=-1) < 0.99))


def pack_emit_np(emit):
    """list of dict(mu, sigma, kappa, nu) -> (K, plen) float64 in the layout of include/svihmm.h."""
    rows = []
    for e in emit:
        if "alpha" in e:
            rows.append(np.asarray(e["alpha"], dtype=np.float64).ravel())
            continue
        mu = np.asarray(e["mu"], dtype=np.float64).ravel()
        D = mu.size
        sg = np.asarray(e["sigma"], dtype=np.float64)
        if sg.ndim == 2:
            rows.append(np.concatenate([mu, sg.ravel(), [float(e["kappa"])], [float(e["nu"])]]))
        else:
            rows.append(np.concatenate([mu, sg, np.broadcast_to(np.asarray(e["kappa"], float), (D,)),
                                        np.broadcast_to(np.asarray(e["nu"], float), (D,))]))
    return np.array(rows)


def make_random_problem(seed, K, D, T_full, kind="niw_full", miss=0.0, sep=0.4):
    """Seeded synthetic problem with OVERLAPPING states (mean spread `sep` sigma) so that the
    posteriors are not one-hot (vacuous-parity trap).  Sticky chain as in
    SURVEY section 8d.  Returns obs, mask, var_tran, emit, prior_tran, prior_emit."""
    rs = np.random.RandomState(seed)
    tran = 0.9 * np.eye(K) + 0.1 / max(K - 1, 1) * (1 - np.eye(K)) if K > 1 else np.ones((1, 1))
    mus = sep * rs.randn(K, D)
    sts = np.empty(T_full, dtype=np.int64)
    st = 0
    u = rs.rand(T_full)
    cdf = np.cumsum(tran, axis=1)
    for t in range(T_full):
        sts[t] = st
        st = min(int(np.searchsorted(cdf[st], u[t])), K - 1)
    obs = mus[sts] + rs.randn(T_full, D)
    mask = rs.rand(T_full) < miss
    emit, prior_emit = [], []
    for k in range(K):
        mu = mus[k] + 0.3 * rs.randn(D)
        if kind == "niw_full":
            A = rs.randn(D, D) * 0.2
            nu = D + 3. + 2 * rs.rand()
            emit.append(dict(mu=mu, sigma=(nu - D - 1) * (np.eye(D) + A.dot(A.T)), kappa=0.7 + rs.rand(), nu=nu))
            prior_emit.append(dict(mu=np.zeros(D), sigma=0.75 * np.eye(D), kappa=0.01, nu=D + 2.))
        else:
            nu = 4. + 2 * rs.rand(D)
            emit.append(dict(mu=mu, sigma=(nu - 2) * (1. + 0.3 * rs.rand(D)), kappa=0.7 + rs.rand(D), nu=nu))
            prior_emit.append(dict(mu=np.zeros(D), sigma=0.75 * np.ones(D), kappa=0.01 * np.ones(D),
                                   nu=3. * np.ones(D)))
    return dict(obs=obs, sts=sts, mask=mask, var_tran=1. + 5. * rs.rand(K, K), emit=emit,
                prior_tran=np.ones((K, K)), prior_emit=prior_emit)


def make_categorical_problem(seed, K, C, T_full, miss=0.0):
    """Seeded categorical-emission problem with overlapping symbol distributions."""
    rs = np.random.RandomState(seed)
    tran = 0.9 * np.eye(K) + 0.1 / max(K - 1, 1) * (1 - np.eye(K))
    pmf = rs.dirichlet(2. * np.ones(C), K)
    sts = np.empty(T_full, dtype=np.int64)
    st = 0
    u = rs.rand(T_full)
    cdf = np.cumsum(tran, axis=1)
    for t in range(T_full):
        sts[t] = st
        st = min(int(np.searchsorted(cdf[st], u[t])), K - 1)
    cum = np.cumsum(pmf, axis=1)
    obs = np.minimum((rs.rand(T_full)[:, None] > cum[sts]).sum(1), C - 1).astype(np.float64)[:, None]
    mask = rs.rand(T_full) < miss
    emit = [dict(alpha=0.5 + 20. * pmf[k] + rs.rand(C)) for k in range(K)]
    prior_emit = [dict(alpha=1. + 0.5 * rs.rand(C)) for _ in range(K)]
    return dict(obs=obs, sts=sts, mask=mask, var_tran=1. + 5. * rs.rand(K, K), emit=emit,
                prior_tran=np.ones((K, K)), prior_emit=prior_emit)
