"""Parity of the CUDA path (through the C ABI of libsvihmm.so) against
 (a) outputs of the reference itself (tests/golden/*.npz), and
 (b) the float64 oracle on seeded inputs at sizes the oracle finishes in seconds.

Tolerances (north_star: state marginals within 1e-5 relative of the reference):
  var_x        |q - q_ref| <= 1e-5 * q_ref + 2e-7       (float32 output; 2e-7 ~ 2 ulp of 1.0)
  statistics   rtol 1e-5 relative to the largest entry of each block
  new globals  rtol 1e-5 (sigma: relative to its largest entry, eta3 - kappa mu mu^T cancels)
  lliks        float64 kernel: rtol 1e-11
"""
import numpy as np
import pytest

from tests.helpers import (SVI_CASES, emit_list, frac_soft, golden_prior_emit, load_golden,
                           make_categorical_problem, make_random_problem, pack_emit_np)

pytestmark = pytest.mark.gpu

Q_RTOL, Q_ATOL = 1e-5, 2e-7
S_RTOL = 1e-5


def _engine(K, D, kind="niw_full"):
    from pysvihmm_b200.engine import EStepEngine
    return EStepEngine(K, D, kind)


def assert_q(q, q_ref):
    err = np.abs(q.astype(np.float64) - q_ref)
    bound = Q_RTOL * q_ref + Q_ATOL
    worst = float(np.max(err - bound))
    assert worst <= 0, "marginals off: max excess %.3e, max abs err %.3e" % (worst, err.max())


def assert_block(a, ref, rtol=S_RTOL, what=""):
    scale = max(float(np.max(np.abs(ref))), 1e-30)
    err = float(np.max(np.abs(np.asarray(a, dtype=np.float64) - ref))) / scale
    assert err <= rtol, "%s: relative-to-max error %.3e > %.1e" % (what, err, rtol)


def oracle_stats(O, r, obs, mask, starts, T, wrap):
    """Summed (over the minibatch) statistics from oracle marginals."""
    K = r["var_x"].shape[-1]
    A = O.tran_stat(r["var_x"], wrap).sum(0)
    n = np.zeros(K); sx = np.zeros((K, obs.shape[1])); sxx = np.zeros((K, obs.shape[1], obs.shape[1]))
    for b, s in enumerate(starts):
        inds = ~mask[s:s + T]
        x = np.nan_to_num(obs[s:s + T][inds])
        w = r["var_x"][b][inds]
        n += w.sum(0)
        sx += w.T.dot(x)
        sxx += np.einsum("tk,ti,tj->kij", w, x, x)
    return A, n, sx, sxx


# "fused": the one-CTA-per-window kernels (fused.cuh / fused_pipe.cuh: the engine's choice for the small
# minibatches of these cases); "batched": the batched tensor-core path (batch16.cuh) forced for every
# eligible call; "unfused": per-phase kernels (KEEP_LOCALS)
PATHS = ["fused", "batched", "unfused"]


def _apply_path(eng, path):
    from pysvihmm_b200 import _lib as L
    eng.set_tuning(L.TUNE_B16_MIN_B, 1 if path == "batched" else 0)


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("obs_dtype", ["f64", "f32"])
@pytest.mark.parametrize("name", SVI_CASES)
def test_svi_step_matches_reference_golden(name, obs_dtype, path):
    """One engine call per global step reproduces the reference's per-window tables, summed
    statistics (quirks Q1/Q2/Q5) and natural-gradient update (hmmsgd_metaobs.py:405-439)."""
    from pysvihmm_b200 import _lib as L
    g = load_golden(name)
    obs, mask = g["obs"], g["mask"]
    Lh, S = int(g["L"]), int(g["mb_sz"])
    T = 2 * Lh + 1
    K, D = g["init_tran"].shape[0], obs.shape[1]
    if path == "batched":
        pytest.skip("full-covariance fixtures never take the batched path")
    eng = _engine(K, D)
    _apply_path(eng, path)
    eng.set_series(obs, mask, dtype=obs_dtype)
    pe = golden_prior_emit(g, K)
    eng.set_prior(g["prior_tran"], pack_emit_np(pe))
    eng.set_globals(g["init_tran"], pack_emit_np(emit_list(g["init_mu"], g["init_sigma"],
                                                            g["init_kappa"], g["init_nu"])))
    f32 = obs_dtype == "f32"
    for it in range(int(g["maxit"])):
        starts = g["w_starts"][it]
        unf = path == "unfused"
        vx, stats = eng.estep(starts, T, flags=L.WRAP | L.ADD_PRIOR, keep_locals=unf)
        q = vx.cpu().numpy()
        # float32 storage of the series perturbs ll by ~|x| * 6e-8 * |dll/dx|: not a kernel error
        if f32:
            assert np.max(np.abs(q - g["w_var_x"][it])) < 2e-4
        else:
            assert_q(q, g["w_var_x"][it])
            loc = eng.get_locals(S, T, tables=unf)
            la = g["w_lalpha"][it]
            np.testing.assert_allclose(loc["logZ"], np.logaddexp.reduce(la[:, -1], axis=-1), rtol=2e-6)
            np.testing.assert_allclose(loc["lb_q4"], g["w_lb"][it], rtol=2e-6)
            if unf:
                np.testing.assert_allclose(loc["lliks"], g["w_ll"][it], rtol=1e-11, atol=1e-11)
                # softmax(lalpha[t]) == normalised forward message
                sm = np.exp(la - np.logaddexp.reduce(la, axis=-1)[..., None])
                assert np.max(np.abs(loc["alpha"] - sm)) < 2e-6
        s = eng.unpack_stats(stats)
        tol = 2e-4 if f32 else S_RTOL
        assert_block(s["A"], g["w_A_i"][it].sum(0), tol, "A_inter")
        assert_block(s["n"], g["w_e2"][it].sum(0), tol, "n")
        assert_block(s["sx"], g["w_e1"][it].sum(0), tol, "sx")
        assert_block(s["sxx"], g["w_e3"][it].sum(0), tol, "sxx")
        assert s["B"] == S
        bA = (obs.shape[0] - 2 * Lh - 1) / (2. * Lh * S)
        bE = (obs.shape[0] - 2 * Lh - 1) / ((2. * Lh + 1.) * S)
        eng.global_update(stats, float(g["g_lrate"][it]), bA, bE)
        vt, vi, em = eng.get_globals()
        e = eng.unpack_emit(em)
        gtol = 5e-4 if f32 else S_RTOL
        assert_block(vt, g["g_var_tran"][it], gtol, "var_tran")
        assert_block(e["mu"], g["g_mu"][it], gtol, "mu")
        assert_block(e["sigma"], g["g_sigma"][it], gtol, "sigma")
        assert_block(e["kappa"], g["g_kappa"][it], gtol, "kappa")
        assert_block(e["nu"], g["g_nu"][it], gtol, "nu")
        if it + 1 < int(g["maxit"]):
            np.testing.assert_allclose(vi, g["w_var_init"][it + 1][0], rtol=1e-5, atol=1e-9)
        if not f32:
            # keep the next iteration on the reference's trajectory (isolates per-step error)
            eng.set_globals(g["g_var_tran"][it], pack_emit_np(emit_list(
                g["g_mu"][it], g["g_sigma"][it], g["g_kappa"][it], g["g_nu"][it])))
    eng.close()


def test_stationary_init_matches_reference():
    """Quirk Q3 (hmmsgd_metaobs.py:413-418): |top eigenvector|, unit L2 norm."""
    g = load_golden("svi_k16_d8_l50")
    K, D = 16, 8
    eng = _engine(K, D)
    eng.set_globals(g["init_tran"], pack_emit_np(emit_list(g["init_mu"], g["init_sigma"],
                                                            g["init_kappa"], g["init_nu"])))
    _, vi, _ = eng.get_globals()
    np.testing.assert_allclose(vi, g["w_var_init"][0][0], rtol=1e-9)
    assert abs(np.linalg.norm(vi) - 1.0) < 1e-12
    eng.close()


@pytest.mark.parametrize("K", [40, 64, 100, 256, 300])
def test_stationary_init_large_K_matches_oracle(K):
    """The same vector for K > 32: one-CTA GTH elimination (K <= 64) and the cluster kernel with the matrix in
    distributed shared memory (64 < K <= 384, gth_cluster.cuh) against the oracle's eig-based restatement of
    hmmsgd_metaobs.py:413-418, for a sticky and for a nearly uniform transition matrix."""
    from oracle import svihmm_oracle as O
    rs = np.random.RandomState(K)
    D = 2
    emit = emit_list(rs.randn(K, D), np.tile(np.eye(D)[None], (K, 1, 1)), np.ones(K), (D + 3.) * np.ones(K))
    for var_tran in (1. + 5. * rs.rand(K, K), 0.05 + rs.rand(K, K) + 40. * np.eye(K)):
        eng = _engine(K, D)
        eng.set_globals(var_tran, pack_emit_np(emit))
        _, vi, _ = eng.get_globals()
        np.testing.assert_allclose(vi, O.stationary_init(var_tran), rtol=1e-9)
        assert abs(np.linalg.norm(vi) - 1.0) < 1e-12
        eng.close()


def test_batch_cavi_matches_reference_golden():
    """hmmbatchcd.VBHMM.infer iterations (hmmbatchcd.py:135-141,172-189) on device."""
    g = load_golden("cavi_k2_d2_t200")
    K, D, T = 2, 2, g["obs"].shape[0]
    eng = _engine(K, D)
    eng.set_series(g["obs"], g["mask"], dtype="f64")
    eng.set_prior(g["prior_tran"], pack_emit_np(golden_prior_emit(g, K)), g["prior_init"])
    eng.set_globals(g["init_var_tran"], pack_emit_np(emit_list(
        g["init_mu"], g["init_sigma"], g["init_kappa"], g["init_nu"])), g["init_var_init"])
    for it in range(len(g["it_lZ"])):
        vx, stats = eng.estep([0], T, flags=0)
        assert_q(vx[0].cpu().numpy(), g["it_var_x"][it])
        s = eng.unpack_stats(stats)
        np.testing.assert_allclose(s["lb_q4"], g["it_lZ"][it], rtol=2e-6)
        eng.batch_update(stats)
        vt, vi, em = eng.get_globals()
        e = eng.unpack_emit(em)
        assert_block(vi, g["it_var_init"][it], S_RTOL, "var_init")
        assert_block(vt, g["it_var_tran"][it], S_RTOL, "var_tran")
        assert_block(e["mu"], g["it_mu"][it], S_RTOL, "mu")
        assert_block(e["sigma"], g["it_sigma"][it], S_RTOL, "sigma")
        assert_block(e["kappa"], g["it_kappa"][it], S_RTOL, "kappa")
        assert_block(e["nu"], g["it_nu"][it], S_RTOL, "nu")
    eng.close()


# (K, D, T, B, kind): covers every kernel mapping (lane groups KP=2..32, wide K>32), odd sizes
ORACLE_CASES = [
    (2, 1, 7, 3, "niw_full"),
    (3, 2, 1, 4, "niw_full"),          # T = 1: no transitions at all
    (7, 5, 33, 9, "niw_full"),
    (16, 8, 512, 12, "niw_diag"),      # BASELINE config 2 shape (reduced B)
    (16, 8, 511, 6, "niw_full"),
    (32, 16, 129, 5, "niw_full"),
    (33, 4, 64, 5, "niw_full"),        # first wide size
    (64, 32, 257, 4, "niw_full"),      # BASELINE config 3 shape (reduced B, T)
    (64, 32, 1024, 2, "niw_full"),     # BASELINE config 3 window (full T, reduced B)
    (100, 3, 50, 3, "niw_diag"),
    (256, 64, 256, 2, "niw_full"),     # BASELINE config 4 shape (reduced B; float32 recursions, not bf16)
    (300, 4, 40, 2, "niw_diag"),       # K*K floats exceed shared memory: transition matrix read through L2
    (32, 4, 100, 3, "niw_diag"),       # fused path, one chain per warp (KP = 32)
    (20, 3, 300, 4, "niw_full"),       # fused path, K not a power of two
    (4, 12, 40, 3, "niw_diag"),        # fused path, D > K: observation staging in several tiles
    (5, 9, 23, 2, "niw_full"),
    (2, 2, 20000, 1, "niw_full"),      # too long for shared memory: per-phase kernels either way
    (16, 8, 257, 37, "niw_diag"),      # batched path: 3 groups of 16 windows (last one partial), T % 8 = 1
    (11, 5, 40, 17, "niw_diag"),       # batched path: K, D not multiples of anything, one window in the 2nd group
    (3, 13, 9, 2, "niw_diag"),         # batched path: D > 8 (two feature tiles), T barely over one unit
    (16, 16, 1, 5, "niw_diag"),        # batched path: T = 1
    (2, 1, 3000, 3, "niw_diag"),       # batched path: long windows
]


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("K,D,T,B,kind", ORACLE_CASES)
def test_estep_matches_oracle(K, D, T, B, kind, path):
    from oracle import svihmm_oracle as O
    from pysvihmm_b200 import _lib as L
    if path == "batched" and not (kind == "niw_diag" and K <= 16 and D <= 16):
        pytest.skip("the batched path does not take this shape")
    big = K >= 64            # log-domain recursions cost B*K*K logaddexp per step: use the pinned scaled form
    p = make_random_problem(seed=K * 1000 + T, K=K, D=D, T_full=max(4 * T, 300), kind=kind, miss=0.1)
    starts = np.random.RandomState(5).randint(0, p["obs"].shape[0] - T + 1, B)
    eng = _engine(K, D, kind)
    _apply_path(eng, path)
    eng.set_series(p["obs"], p["mask"], dtype="f64")
    eng.set_prior(p["prior_tran"], pack_emit_np(p["prior_emit"]))
    eng.set_globals(p["var_tran"], pack_emit_np(p["emit"]))
    for flags, wrap, mask_ll in [(L.WRAP | L.ADD_PRIOR, True, False), (L.MASK_LL, False, True)]:
        unf = path == "unfused"
        vx, stats = eng.estep(starts, T, flags=flags, keep_locals=unf)
        r = O.svi_minibatch_step(p["obs"], p["mask"], starts, T, p["var_tran"], p["emit"],
                                 p["prior_tran"], p["prior_emit"], 0.5, max(T // 2, 1), wrap=wrap,
                                 mask_ll=mask_ll, scaled=big)
        q = vx.cpu().numpy()
        assert frac_soft(r["var_x"]) > 0.2, "vacuous parity: posteriors are one-hot"
        assert_q(q, r["var_x"])
        loc = eng.get_locals(B, T, tables=unf)
        if unf:
            np.testing.assert_allclose(loc["lliks"], r["ll"], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(loc["logZ"], r["logZ"], rtol=3e-6, atol=1e-4)
        s = eng.unpack_stats(stats)
        A = np.zeros((K, K)); n = np.zeros(K); sx = np.zeros((K, D))
        sxx = np.zeros((K, D, D) if kind == "niw_full" else (K, D))
        A = O.tran_stat(r["var_x"], wrap).sum(0) + (B * (p["prior_tran"] - 1.) if wrap else 0.)
        for k in range(K):
            n[k] = r["emit_inter"][k][1]
            sx[k] = r["emit_inter"][k][0]
            sxx[k] = r["emit_inter"][k][2]
        assert_block(s["A"], A, S_RTOL, "A")
        assert_block(s["n"], n, S_RTOL, "n")
        assert_block(s["sx"], sx, S_RTOL, "sx")
        assert_block(s["sxx"], sxx, S_RTOL, "sxx")
        assert_block(s["q0"], r["var_x"][:, 0].sum(0), S_RTOL, "q0")
        np.testing.assert_allclose(s["logZ"], r["logZ"].sum(), rtol=3e-6)
        np.testing.assert_allclose(s["lb_q4"], r["lb"], rtol=3e-6)
    # natural-gradient step from the (WRAP|ADD_PRIOR) statistics
    vx, stats = eng.estep(starts, T, flags=L.WRAP | L.ADD_PRIOR)
    r = O.svi_minibatch_step(p["obs"], p["mask"], starts, T, p["var_tran"], p["emit"],
                             p["prior_tran"], p["prior_emit"], 0.37, max(T // 2, 1), wrap=True, scaled=big)
    Lh, Tf = max(T // 2, 1), p["obs"].shape[0]
    eng.global_update(stats, 0.37, (Tf - 2 * Lh - 1) / (2. * Lh * B), (Tf - 2 * Lh - 1) / ((2. * Lh + 1.) * B))
    vt, vi, em = eng.get_globals()
    e = eng.unpack_emit(em)
    assert_block(vt, r["var_tran_new"], S_RTOL, "var_tran")
    for key in ("mu", "sigma", "kappa", "nu"):
        ref = np.array([np.broadcast_to(x[key], e[key][0].shape) for x in r["emit_new"]])
        assert_block(e[key], ref, S_RTOL, key)
    np.testing.assert_allclose(vi, O.stationary_init(r["var_tran_new"]), rtol=1e-5, atol=1e-9)
    eng.close()


@pytest.mark.parametrize("K,C,T,B", [(4, 6, 50, 5), (16, 20, 200, 7), (40, 9, 64, 3)])
def test_categorical_emissions_match_oracle(K, C, T, B):
    """SVIHMM_EMIT_CATEGORICAL (Categorical.expected_log_likelihood, distributions.py:1383-1386; count
    statistics and Dirichlet natural-gradient step, hmmsgd_metaobs.py:907-926,1071-1084) against the
    oracle; the expected log-likelihood table also against the reference's own values (cat_ell)."""
    from oracle import svihmm_oracle as O
    from pysvihmm_b200 import _lib as L
    p = make_categorical_problem(seed=K + C, K=K, C=C, T_full=max(4 * T, 300), miss=0.1)
    obs = p["obs"].copy()
    obs[np.random.RandomState(2).rand(len(obs)) < 0.03] = np.nan        # NaN symbols = missing
    starts = np.random.RandomState(5).randint(0, obs.shape[0] - T + 1, B)
    eng = _engine(K, C, "categorical")
    eng.set_series(obs, p["mask"], dtype="f64")
    eng.set_prior(p["prior_tran"], pack_emit_np(p["prior_emit"]))
    eng.set_globals(p["var_tran"], pack_emit_np(p["emit"]))
    vx, stats = eng.estep(starts, T, flags=L.WRAP | L.ADD_PRIOR, keep_locals=True)
    Lh, Tf = max(T // 2, 1), obs.shape[0]
    r = O.svi_minibatch_step(obs, p["mask"], starts, T, p["var_tran"], p["emit"], p["prior_tran"],
                             p["prior_emit"], 0.37, Lh, wrap=True)
    assert frac_soft(r["var_x"]) > 0.2, "vacuous parity: posteriors are one-hot"
    loc = eng.get_locals(B, T)
    np.testing.assert_allclose(loc["lliks"], r["ll"], rtol=1e-11, atol=1e-12)
    assert_q(vx.cpu().numpy(), r["var_x"])
    s = eng.unpack_stats(stats)
    assert_block(s["A"], r["A_inter"], S_RTOL, "A")
    assert_block(s["sx"], r["counts"], S_RTOL, "counts")
    assert_block(s["n"], r["counts"].sum(1), S_RTOL, "n")
    np.testing.assert_allclose(s["lb_q4"], r["lb"], rtol=3e-6)
    eng.global_update(stats, 0.37, (Tf - 2 * Lh - 1) / (2. * Lh * B), (Tf - 2 * Lh - 1) / ((2. * Lh + 1.) * B))
    vt, vi, em = eng.get_globals()
    assert_block(vt, r["var_tran_new"], S_RTOL, "var_tran")
    assert_block(em, np.array([e["alpha"] for e in r["emit_new"]]), S_RTOL, "alpha")
    # second E-step on the updated globals exercises the refreshed log-probability table
    vx2, _ = eng.estep(starts, T, flags=L.WRAP | L.ADD_PRIOR)
    r2 = O.svi_minibatch_step(obs, p["mask"], starts, T, r["var_tran_new"], r["emit_new"], p["prior_tran"],
                              p["prior_emit"], 0.37, Lh, wrap=True)
    assert_q(vx2.cpu().numpy(), r2["var_x"])
    eng.close()


def test_categorical_ell_table_matches_reference_golden():
    g = load_golden("cat_ell")
    K, C = g["alpha"].shape
    eng = _engine(K, C, "categorical")
    eng.set_series(g["x"].astype(np.float64)[:, None])
    eng.set_globals(np.ones((K, K)), g["alpha"])
    eng.estep([0], len(g["x"]), want_var_x=False, keep_locals=True)
    ll = eng.get_locals(1, len(g["x"]))["lliks"][0]
    np.testing.assert_allclose(ll.T, g["ell"], rtol=1e-12, atol=1e-13)
    eng.close()


@pytest.mark.parametrize("K,C,D,T,B,kind", [(3, 2, 2, 40, 4, "niw_full"), (8, 4, 16, 128, 3, "niw_diag"),
                                             (32, 4, 16, 96, 2, "niw_full"), (40, 3, 5, 64, 3, "niw_full")])
def test_gmm_emissions_match_oracle(K, C, D, T, B, kind):
    """EXTENSION (BASELINE config 5): mixture-of-NIW emissions per state.  No reference mean-field
    code exists (parity unpinned); the oracle restates labels.py:52-65 + distributions.py:351-366 and
    reduces to the pinned Gaussian path at C = 1 (test_oracle_golden)."""
    from oracle import svihmm_oracle as O
    from pysvihmm_b200 import _lib as L
    p = make_random_problem(seed=K * 100 + C, K=K * C, D=D, T_full=max(4 * T, 300), kind=kind, miss=0.1, sep=0.8)
    rs = np.random.RandomState(K + C)
    emit = [dict(omega=1. + 3. * rs.rand(C), comps=p["emit"][k * C:(k + 1) * C]) for k in range(K)]
    prior = [dict(omega=0.5 + rs.rand(C), comps=p["prior_emit"][k * C:(k + 1) * C]) for k in range(K)]
    var_tran, prior_tran = 1. + 5. * rs.rand(K, K), np.ones((K, K))
    obs = p["obs"].copy()
    obs[rs.rand(len(obs)) < 0.03, 0] = np.nan
    starts = rs.randint(0, obs.shape[0] - T + 1, B)
    eng = __import__("pysvihmm_b200.engine", fromlist=["EStepEngine"]).EStepEngine(K, D, kind, components=C)
    eng.set_series(obs, p["mask"], dtype="f64")
    eng.set_prior(prior_tran, pack_emit_np(p["prior_emit"]))
    eng.set_mix_weights(np.array([e["omega"] for e in emit]), np.array([e["omega"] for e in prior]))
    eng.set_globals(var_tran, pack_emit_np(p["emit"]))
    Lh, Tf = max(T // 2, 1), obs.shape[0]
    vx, stats = eng.estep(starts, T, flags=L.WRAP | L.ADD_PRIOR)
    r = O.gmm_minibatch_step(obs, p["mask"], starts, T, var_tran, emit, prior_tran, prior, 0.37, Lh)
    assert frac_soft(r["var_x"]) > 0.2, "vacuous parity: posteriors are one-hot"
    assert_q(vx.cpu().numpy(), r["var_x"])
    s = eng.unpack_stats(stats)
    assert_block(s["A"], r["A_inter"], S_RTOL, "A")
    flat = [r["stats"][k][c] for k in range(K) for c in range(C)]
    assert_block(s["n"], np.array([e[1] for e in flat]), S_RTOL, "n")
    assert_block(s["sx"], np.array([e[0] for e in flat]), S_RTOL, "sx")
    assert_block(s["sxx"], np.array([e[2] for e in flat]), S_RTOL, "sxx")
    np.testing.assert_allclose(s["lb_q4"], r["lb"], rtol=3e-6)
    eng.global_update(stats, 0.37, (Tf - 2 * Lh - 1) / (2. * Lh * B), (Tf - 2 * Lh - 1) / ((2. * Lh + 1.) * B))
    vt, vi, em = eng.get_globals()
    e = eng.unpack_emit(em)
    assert_block(vt, r["var_tran_new"], S_RTOL, "var_tran")
    newc = [g for k in range(K) for g in r["emit_new"][k]["comps"]]
    for key in ("mu", "sigma", "kappa", "nu"):
        ref = np.array([np.broadcast_to(x[key], e[key][0].shape) for x in newc])
        assert_block(e[key], ref, S_RTOL, key)
    assert_block(eng.get_mix_weights(), np.array([x["omega"] for x in r["emit_new"]]), S_RTOL, "omega")
    # second E-step: refreshed component constants and expected log-weights
    vx2, _ = eng.estep(starts, T, flags=L.WRAP | L.ADD_PRIOR)
    r2 = O.gmm_minibatch_step(obs, p["mask"], starts, T, r["var_tran_new"], r["emit_new"], prior_tran, prior, 0.37, Lh)
    assert_q(vx2.cpu().numpy(), r2["var_x"])
    eng.close()


@pytest.mark.parametrize("K,D,T,B", [(256, 8, 64, 130), (128, 4, 40, 7), (96, 4, 33, 129)])
def test_bf16_dense_tensor_core_step(K, D, T, B):
    """SVIHMM_BF16_DENSE (BASELINE config 4: the K x K step as a dense contraction on tcgen05 tensor
    cores, bf16 messages, float32 accumulators).  The reference is float64 only, so the tolerance is
    restated: |q - q_ref| <= 3e-2 absolute on the marginals (bf16 has 8 mantissa bits; errors do not
    accumulate because the recursions forget), statistics 3e-2 relative to the largest entry, logZ 1e-2
    relative.  The same inputs WITHOUT the flag meet the 1e-5 bound (test_estep_matches_oracle)."""
    from oracle import svihmm_oracle as O
    from pysvihmm_b200 import _lib as L
    p = make_random_problem(seed=K + T, K=K, D=D, T_full=max(6 * T, 400), kind="niw_diag", miss=0.05, sep=1.5)
    starts = np.random.RandomState(5).randint(0, p["obs"].shape[0] - T + 1, B)
    eng = _engine(K, D, "niw_diag")
    eng.set_series(p["obs"], p["mask"], dtype="f64")
    eng.set_prior(p["prior_tran"], pack_emit_np(p["prior_emit"]))
    eng.set_globals(p["var_tran"], pack_emit_np(p["emit"]))
    vx, stats = eng.estep(starts, T, flags=L.WRAP | L.ADD_PRIOR | L.BF16_DENSE)
    nb = min(B, 8)                                        # oracle on a few windows of each CTA
    pick = np.r_[0:nb // 2, B - nb // 2:B]
    r = O.svi_minibatch_step(p["obs"], p["mask"], starts[pick], T, p["var_tran"], p["emit"], p["prior_tran"],
                             p["prior_emit"], 0.5, max(T // 2, 1))
    q = vx.cpu().numpy()[pick]
    assert np.isfinite(q).all() and np.allclose(q.sum(-1), 1., atol=1e-4)
    assert float(np.max(np.abs(q - r["var_x"]))) < 3e-2
    vx32, stats32 = eng.estep(starts, T, flags=L.WRAP | L.ADD_PRIOR)          # float32 recursions
    s16, s32 = eng.unpack_stats(stats), eng.unpack_stats(stats32)
    for key in ("A", "n", "sx", "sxx"):
        assert_block(s16[key], s32[key], 3e-2, key)
    assert abs(s16["logZ"] - s32["logZ"]) < 1e-2 * abs(s32["logZ"])
    eng.close()


@pytest.mark.parametrize("flags_extra", ["wrap", "nowrap"])
def test_bf16_dense_statistics_many_pairs_per_cta(flags_extra):
    """The tcgen05 statistics kernels with MORE (tile, t) steps than CTAs, so that every CTA accumulates
    several pairs in tensor memory (operand ring reuse, accumulate flag, mbarrier phases): 3 tiles of
    windows and T = 200 give ~4 pairs per CTA; with and without the wrap-around pair (Q2).  Checked
    against the float32 statistics computed by k_stats from the SAME bf16-path marginals (3e-3 relative
    to the largest entry: only the bf16 rounding of q separates the two) and against the float32 path."""
    from pysvihmm_b200 import _lib as L
    K, D, T, B = 128, 4, 200, 300
    p = make_random_problem(seed=77, K=K, D=D, T_full=2400, kind="niw_diag", miss=0.05, sep=1.5)
    starts = np.random.RandomState(6).randint(0, p["obs"].shape[0] - T + 1, B)
    eng = _engine(K, D, "niw_diag")
    eng.set_series(p["obs"], p["mask"], dtype="f64")
    eng.set_prior(p["prior_tran"], pack_emit_np(p["prior_emit"]))
    eng.set_globals(p["var_tran"], pack_emit_np(p["emit"]))
    fl = L.ADD_PRIOR | (L.WRAP if flags_extra == "wrap" else 0)
    vx, stats = eng.estep(starts, T, flags=fl | L.BF16_DENSE)
    s16 = eng.unpack_stats(stats)
    # float64 statistics from the marginals the dense path returned (product-of-marginals rule, Q1/Q2)
    q = vx.cpu().numpy().astype(np.float64)
    A = np.einsum("wti,wtj->ij", q[:, :-1], q[:, 1:])
    if flags_extra == "wrap":
        A += np.einsum("wi,wj->ij", q[:, -1], q[:, 0])
    A += B * (p["prior_tran"] - 1.)
    assert_block(s16["A"], A, 6e-3, "A from returned marginals")
    idx = starts[:, None] + np.arange(T)[None, :]
    x = np.nan_to_num(p["obs"][idx]); w = (~p["mask"][idx].astype(bool)) & ~np.isnan(p["obs"][idx]).any(-1)
    qw = q * w[..., None]
    assert_block(s16["n"], qw.sum((0, 1)), 6e-3, "n")
    assert_block(s16["sx"], np.einsum("wtk,wtd->kd", qw, x), 6e-3, "sx")
    assert_block(s16["sxx"].reshape(K, -1), np.einsum("wtk,wtd->kd", qw, x * x), 6e-3, "sxx")
    _, stats32 = eng.estep(starts, T, flags=fl)
    s32 = eng.unpack_stats(stats32)
    for key in ("A", "n", "sx", "sxx"):
        assert_block(s16[key], s32[key], 3e-2, key)
    eng.close()


def test_exact_xi_option_matches_oracle():
    """SVIHMM_EXACT_XI (not reference behaviour): sum_t of the true pairwise posteriors."""
    from oracle import svihmm_oracle as O
    from pysvihmm_b200 import _lib as L
    for K, D, T, B in [(5, 3, 40, 6), (40, 4, 30, 3)]:
        p = make_random_problem(seed=77 + K, K=K, D=D, T_full=400, kind="niw_full", miss=0.0)
        starts = np.arange(B) * 17
        eng = _engine(K, D)
        eng.set_series(p["obs"], None, dtype="f64")
        eng.set_globals(p["var_tran"], pack_emit_np(p["emit"]))
        _, stats = eng.estep(starts, T, flags=L.EXACT_XI)
        xw = p["obs"][starts[:, None] + np.arange(T)[None]]
        r = O.local_update(xw, O.stationary_init(p["var_tran"]), p["var_tran"], p["emit"])
        assert_block(eng.unpack_stats(stats)["A"], O.exact_xi_stat(r).sum(0), S_RTOL, "xi")
        eng.close()


def test_nan_rows_carry_no_evidence():
    """np.nan_to_num semantics (hmmsgd_metaobs.py:508-509): a NaN row gives ll[t,:] = 0 and is
    dropped from the emission statistics."""
    from oracle import svihmm_oracle as O
    K, D, T = 4, 3, 60
    p = make_random_problem(seed=3, K=K, D=D, T_full=T, kind="niw_full", miss=0.0)
    obs = p["obs"].copy()
    obs[[0, 7, 8, 30, T - 1]] = np.nan
    obs[12, 1] = np.nan
    eng = _engine(K, D)
    eng.set_series(obs, None, dtype="f64")
    eng.set_globals(p["var_tran"], pack_emit_np(p["emit"]))
    vx, stats = eng.estep([0], T, flags=0)
    r = O.local_update(obs[None], O.stationary_init(p["var_tran"]), p["var_tran"], p["emit"])
    assert np.all(r["ll"][0, 7] == 0)
    assert_q(vx.cpu().numpy(), r["var_x"])
    s = eng.unpack_stats(stats)
    good = ~np.isnan(obs).any(1)
    assert_block(s["n"], r["var_x"][0][good].sum(0), S_RTOL, "n")
    assert_block(s["sx"], r["var_x"][0][good].T.dot(obs[good]), S_RTOL, "sx")
    eng.close()


def test_all_masked_window_and_b1():
    from pysvihmm_b200 import _lib as L
    K, D, T = 3, 2, 20
    p = make_random_problem(seed=4, K=K, D=D, T_full=T, kind="niw_full", miss=0.0)
    eng = _engine(K, D)
    eng.set_series(p["obs"], np.ones(T, bool), dtype="f64")
    eng.set_globals(p["var_tran"], pack_emit_np(p["emit"]))
    vx, stats = eng.estep([0], T, flags=L.MASK_LL)
    s = eng.unpack_stats(stats)
    assert np.all(s["n"] == 0) and np.all(s["sx"] == 0) and np.all(s["sxx"] == 0)
    q = vx.cpu().numpy()
    assert np.allclose(q.sum(-1), 1, atol=1e-6) and np.isfinite(q).all()
    eng.close()


def test_host_call_equals_device_call():
    """svihmm_estep_host (host buffers, windows gathered host->device inside the call) gives the
    same statistics and marginals as svihmm_estep on the HBM-resident series."""
    from pysvihmm_b200 import _lib as L
    K, D, T, B = 16, 8, 128, 37
    for dt in (np.float32, np.float64):
        p = make_random_problem(seed=9, K=K, D=D, T_full=5000, kind="niw_diag", miss=0.05)
        obs = p["obs"].astype(dt)
        starts = np.random.RandomState(1).randint(0, 5000 - T + 1, B)
        eng = _engine(K, D, "niw_diag")
        eng.set_series(obs, p["mask"])
        eng.set_series_streamed(obs, p["mask"])
        eng.set_prior(p["prior_tran"], pack_emit_np(p["prior_emit"]))
        eng.set_globals(p["var_tran"], pack_emit_np(p["emit"]))
        vx, stats = eng.estep(starts, T, flags=L.WRAP | L.ADD_PRIOR)
        vxh, sh = eng.estep_host(starts, T, flags=L.WRAP | L.ADD_PRIOR, want_var_x=True)
        assert np.array_equal(vx.cpu().numpy(), vxh)
        # statistics are accumulated over windows with float64 atomics: order-dependent last bits
        np.testing.assert_allclose(stats.cpu().numpy(), sh, rtol=1e-12, atol=1e-12)
        eng.close()


@pytest.mark.parametrize("D,T,dt", [(8, 300, np.float32), (3, 50, np.float32), (6, 171, np.float64), (32, 129, np.float32)])
def test_window_gather_paths(D, T, dt):
    """The host -> device window gather of svihmm_estep_host: the bulk-copy engine (rows that are 16-byte multiples;
    windows of several 4 KB pieces with a ragged last piece, (8, 300) = 9600 bytes, (32, 129) = 16512 bytes, and a
    float64 series) and the load/store fallback (12-byte rows), with mask bytes: identical marginals, statistics to
    round-off against the HBM-resident series."""
    from pysvihmm_b200 import _lib as L
    K, B = 6, 19
    p = make_random_problem(seed=D * 100 + T, K=K, D=D, T_full=3000, kind="niw_diag", miss=0.1)
    obs = p["obs"].astype(dt)
    starts = np.random.RandomState(2).randint(0, 3000 - T + 1, B)
    starts[0], starts[-1] = 0, 3000 - T
    eng = _engine(K, D, "niw_diag")
    eng.set_series(obs, p["mask"])
    eng.set_series_streamed(obs, p["mask"])
    eng.set_prior(p["prior_tran"], pack_emit_np(p["prior_emit"]))
    eng.set_globals(p["var_tran"], pack_emit_np(p["emit"]))
    vx, stats = eng.estep(starts, T, flags=L.WRAP | L.ADD_PRIOR | L.MASK_LL)
    vxh, sh = eng.estep_host(starts, T, flags=L.WRAP | L.ADD_PRIOR | L.MASK_LL, want_var_x=True)
    assert np.array_equal(vx.cpu().numpy(), vxh)
    np.testing.assert_allclose(stats.cpu().numpy(), sh, rtol=1e-12, atol=1e-12)
    eng.close()


def test_streamed_step_equals_device_step():
    """svihmm_estep_streamed / svihmm_svi_step_host (host series, double-buffered window gather with
    the next minibatch announced one step ahead) follow the same trajectory as svihmm_estep +
    svihmm_global_update on the HBM-resident series, with and without prefetch hits."""
    from pysvihmm_b200 import _lib as L
    K, D, T, B, n = 8, 4, 64, 21, 6
    p = make_random_problem(seed=11, K=K, D=D, T_full=4000, kind="niw_full", miss=0.05)
    obs = p["obs"].astype(np.float32)
    rs = np.random.RandomState(3)
    starts = rs.randint(0, 4000 - T + 1, (n, B))
    flags = L.WRAP | L.ADD_PRIOR

    def fresh():
        eng = _engine(K, D, "niw_full")
        eng.set_prior(p["prior_tran"], pack_emit_np(p["prior_emit"]))
        eng.set_globals(p["var_tran"], pack_emit_np(p["emit"]))
        return eng
    ref = fresh(); ref.set_series(obs, p["mask"])
    ref_stats = []
    for i in range(n):
        _, st = ref.estep(starts[i], T, flags=flags, want_var_x=False)
        ref_stats.append(st.cpu().numpy())
        ref.global_update(st, (i + 1.) ** -0.7, 2.0, 1.5)
    ref_glob = ref.get_globals()
    ref.close()
    # (a) single-call host step; announce the next minibatch except at i == 2 (cold gather at i == 3)
    # and announce a WRONG one at i == 3 (prefetched windows must not be used at i == 4)
    eng = fresh(); eng.set_series_streamed(obs, p["mask"])
    for i in range(n):
        nxt = starts[i + 1] if i + 1 < n and i != 2 else None
        if i == 3:
            nxt = starts[0]
        sh = eng.svi_step_host(starts[i], T, (i + 1.) ** -0.7, 2.0, 1.5, next_starts=nxt, flags=flags)
        np.testing.assert_allclose(sh, ref_stats[i], rtol=1e-9, atol=1e-9)
    for a, b in zip(eng.get_globals(), ref_glob):
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12)
    eng.close()
    # (b) enqueue-only variant with the statistics left on the device
    eng = fresh(); eng.set_series_streamed(obs, p["mask"])
    st = eng.new_stats()
    for i in range(n):
        if i + 2 < n:
            eng.prefetch_windows(starts[i + 2], T)          # two minibatches ahead (ring of 3 slots)
        eng.estep_streamed(starts[i], T, next_starts=starts[i + 1] if i + 1 < n else None, flags=flags, stats=st)
        np.testing.assert_allclose(st.cpu().numpy(), ref_stats[i], rtol=1e-9, atol=1e-9)
        eng.global_update(st, (i + 1.) ** -0.7, 2.0, 1.5)
    for a, b in zip(eng.get_globals(), ref_glob):
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12)
    eng.close()


@pytest.mark.parametrize("K,D,kind", [(16, 8, "niw_diag"), (5, 3, "niw_full")])
def test_svi_run_equals_step_by_step(K, D, kind):
    """svihmm_svi_run (nsteps global steps enqueued by one call, hmmsgd_metaobs.py:396-439) follows the
    trajectory of svihmm_estep + svihmm_global_update called step by step."""
    import torch
    from pysvihmm_b200 import _lib as L
    T, B, n = 64, 40, 5
    p = make_random_problem(seed=21, K=K, D=D, T_full=3000, kind=kind, miss=0.05)
    starts = torch.from_numpy(np.random.RandomState(2).randint(0, 3000 - T + 1, (n, B))).cuda()
    flags = L.WRAP | L.ADD_PRIOR

    def fresh():
        eng = _engine(K, D, kind)
        eng.set_series(p["obs"], p["mask"], dtype="f32")
        eng.set_prior(p["prior_tran"], pack_emit_np(p["prior_emit"]))
        eng.set_globals(p["var_tran"], pack_emit_np(p["emit"]))
        return eng
    a = fresh()
    for i in range(n):
        vxa, st = a.estep(starts[i], T, flags=flags)
        a.global_update(st, (3 + i + 1.0) ** -0.7, 2.0, 1.5)
    b = fresh()
    vxb = torch.empty((B, T, K), dtype=torch.float32, device="cuda")
    stb = b.svi_run(starts, T, 1.0, 0.7, 3, 2.0, 1.5, flags=flags, var_x=vxb)
    torch.cuda.synchronize()
    np.testing.assert_allclose(stb.cpu().numpy(), st.cpu().numpy(), rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(vxb.cpu().numpy(), vxa.cpu().numpy(), rtol=1e-6, atol=1e-9)
    for x, y in zip(a.get_globals(), b.get_globals()):
        np.testing.assert_allclose(x, y, rtol=1e-9, atol=1e-12)
    a.close(); b.close()


@pytest.mark.parametrize("page_lock", [True, False])
def test_memmap_series_streamed_equals_resident(tmp_path, page_lock):
    """gen_synthetic.read_data_mmap's on-disk series (gen_synthetic.py:158-191) as the streamed series:
    read-only mapping page-locked and gathered by the GPU, or (page_lock=False, for files larger than host
    memory) gathered by the CPU into pinned staging; both follow the resident-series trajectory."""
    from pysvihmm_b200 import _lib as L
    K, D, T, B, n, Tf = 6, 4, 48, 13, 4, 3000
    p = make_random_problem(seed=31, K=K, D=D, T_full=Tf, kind="niw_full", miss=0.05)
    fname = str(tmp_path / "obs.dat")
    mm = np.memmap(fname, dtype="float64", mode="w+", shape=(Tf, D))
    mm[:] = p["obs"]; mm.flush(); del mm
    starts = np.random.RandomState(3).randint(0, Tf - T + 1, (n, B))
    flags = L.WRAP | L.ADD_PRIOR

    def fresh():
        eng = _engine(K, D, "niw_full")
        eng.set_prior(p["prior_tran"], pack_emit_np(p["prior_emit"]))
        eng.set_globals(p["var_tran"], pack_emit_np(p["emit"]))
        return eng
    ref = fresh(); ref.set_series(p["obs"], p["mask"], dtype="f64")
    ref_stats = []
    for i in range(n):
        _, st = ref.estep(starts[i], T, flags=flags, want_var_x=False)
        ref_stats.append(st.cpu().numpy())
        ref.global_update(st, (i + 1.) ** -0.7, 2.0, 1.5)
    ref_glob = ref.get_globals(); ref.close()
    eng = fresh()
    eng.set_series_memmap(fname, Tf, D, mask_host=p["mask"], page_lock=page_lock)
    for i in range(n):
        sh = eng.svi_step_host(starts[i], T, (i + 1.) ** -0.7, 2.0, 1.5, next_starts=starts[i + 1] if i + 1 < n else None,
                               flags=flags)
        np.testing.assert_allclose(sh, ref_stats[i], rtol=1e-9, atol=1e-9)
    for a, b in zip(eng.get_globals(), ref_glob):
        np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-12)
    eng.close()


def test_uncentred_series_is_flagged_not_silent():
    """the emission statistics are float32 products, so a series with |mean| >> std loses its
    variance in eta3 - kappa mu mu^T; the global step flags the non-positive scale (svihmm_check /
    get_globals raise) instead of handing back NaN silently.  The same series centred works."""
    from pysvihmm_b200 import SvihmmError, _lib as L
    K, D, T, B = 3, 2, 64, 20
    p = make_random_problem(seed=5, K=K, D=D, T_full=2000, kind="niw_full", miss=0.0, sep=1.0)
    starts = np.random.RandomState(1).randint(0, 2000 - T + 1, B)
    for shift, ok in [(0.0, True), (3e4, False)]:
        eng = _engine(K, D)
        eng.set_series(p["obs"] + shift, None, dtype="f32")
        eng.set_prior(p["prior_tran"], pack_emit_np([dict(e, mu=e["mu"] + shift) for e in p["prior_emit"]]))
        eng.set_globals(p["var_tran"], pack_emit_np([dict(e, mu=e["mu"] + shift) for e in p["emit"]]))
        _, st = eng.estep(starts, T, flags=L.WRAP | L.ADD_PRIOR, want_var_x=False)
        eng.global_update(st, 0.9, 30.0, 30.0)
        if ok:
            eng.get_globals()
        else:
            with pytest.raises(SvihmmError):
                eng.get_globals()
        eng.close()


def test_error_paths():
    from pysvihmm_b200 import SvihmmError
    eng = _engine(3, 2)
    with pytest.raises(SvihmmError):
        eng.estep([0], 5)                      # no globals yet
    p = make_random_problem(seed=1, K=3, D=2, T_full=30, kind="niw_full")
    eng.set_globals(p["var_tran"], pack_emit_np(p["emit"]))
    with pytest.raises(SvihmmError):
        eng.estep([0], 5)                      # no series yet
    eng.set_series(p["obs"], None)
    with pytest.raises(SvihmmError):
        eng.estep([0], 31)                     # window longer than the series
    with pytest.raises(SvihmmError):
        eng.estep([0], 5, flags=2)             # ADD_PRIOR without priors
    eng.close()
    with pytest.raises(SvihmmError):
        _engine(0, 2)
