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
