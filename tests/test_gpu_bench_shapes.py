"""Parity of exactly the code paths bench.py times, at the BASELINE shapes and dtypes, against the
float64 oracle (scaled-domain form, held to the reference-made fixtures in test_oracle_golden).

  c2  K16 D8 T512, float32 series, B = 256 and B > CTAs in flight   -> k_estep_pipe (vecx branch) /
                                                                        the batched tensor-core path
  c3  K64 D32 T1024 full covariance, float32 series, B = 64          -> k_emit_tc + wide64 kernels + k_stats_tc
  c4  K256 D64 T256 B = 256, SVIHMM_BF16_DENSE                        -> tcgen05 recursion + statistics
  c5  K32 x 4 components D16, T = 256 and T = 2048                    -> mixture path

A float32 series is the benchmark's input: the oracle then sees obs.astype(float32).astype(float64),
i.e. the SAME numbers, and the 1e-5 bound applies unchanged (tolerances as in test_gpu_parity).
"""
import numpy as np
import pytest

from tests.helpers import frac_soft, make_random_problem, pack_emit_np
from tests.test_gpu_parity import Q_ATOL, Q_RTOL, S_RTOL, assert_block, assert_q

pytestmark = pytest.mark.gpu


def _stats_from_q(q, obs, mask, starts, T, wrap, prior_tran, full):
    """Summed minibatch statistics (Q1 product of marginals, Q2 wrap, Q5 prior per window;
    hmmsgd_metaobs.py:873-904) from oracle marginals q (B,T,K), vectorised over the minibatch."""
    B, _, K = q.shape
    idx = np.asarray(starts)[:, None] + np.arange(T)[None]
    x = obs[idx]
    keep = ~(mask[idx].astype(bool)) & ~np.isnan(x).any(-1)
    x = np.nan_to_num(x)
    A = np.einsum("bti,btj->ij", q[:, :-1], q[:, 1:])
    if wrap:
        A += np.einsum("bi,bj->ij", q[:, -1], q[:, 0])
        A += B * (prior_tran - 1.)
    w = q * keep[..., None]
    n = w.sum((0, 1))
    sx = np.einsum("btk,btd->kd", w, x)
    sxx = np.einsum("btk,bti,btj->kij", w, x, x) if full else np.einsum("btk,btd->kd", w, x * x)
    return A, n, sx, sxx


def _run_case(K, D, T, B, kind, dtype, flags_extra=0, q_check=None, seed_off=0, sep=0.4, miss=0.05,
              q_abs=None, s_rtol=S_RTOL, lz_rtol=3e-6, onecta=False, nan_frac=0.0):
    from oracle import svihmm_oracle as O
    from pysvihmm_b200 import _lib as L
    from pysvihmm_b200.engine import EStepEngine
    p = make_random_problem(seed=K * 1000 + T + seed_off, K=K, D=D, T_full=max(8 * T, 4000), kind=kind,
                            miss=miss, sep=sep)
    obs = p["obs"]
    if nan_frac > 0:                                             # NaN entries: those rows carry no evidence
        obs = obs.copy()
        rsn = np.random.RandomState(77)
        obs[rsn.rand(obs.shape[0]) < nan_frac, rsn.randint(0, D)] = np.nan
        # the reference drops only MASKED rows from the emission statistics (a NaN observation would
        # poison its sums, hmmsgd_metaobs.py:884-904); flag the NaN rows as masked so that the oracle's
        # update stays finite - the engine drops them either way
        p["mask"] = p["mask"] | np.isnan(obs).any(1)
    if dtype == "f32":
        obs = obs.astype(np.float32).astype(np.float64)          # the numbers the engine is given
    starts = np.random.RandomState(5).randint(0, obs.shape[0] - T + 1, B)
    eng = EStepEngine(K, D, kind)
    # the one-CTA-per-window kernels, or the batched tensor-core path forced (the engine's own choice
    # switches between them at the measured crossover, B = 4096)
    eng.set_tuning(L.TUNE_B16_MIN_B, 0 if onecta else 1)
    eng.set_series(obs, p["mask"], dtype=dtype)
    eng.set_prior(p["prior_tran"], pack_emit_np(p["prior_emit"]))
    eng.set_globals(p["var_tran"], pack_emit_np(p["emit"]))
    fl = L.WRAP | L.ADD_PRIOR | flags_extra
    vx, stats = eng.estep(starts, T, flags=fl)
    r = O.svi_minibatch_step(obs, p["mask"], starts, T, p["var_tran"], p["emit"], p["prior_tran"],
                             p["prior_emit"], 0.37, max(T // 2, 1), wrap=True, scaled=True)
    q = vx.cpu().numpy()
    assert frac_soft(r["var_x"]) > 0.2, "vacuous parity: posteriors are one-hot"
    if q_abs is None:
        assert_q(q, r["var_x"])
    else:
        assert np.isfinite(q).all() and float(np.max(np.abs(q - r["var_x"]))) < q_abs
    s = eng.unpack_stats(stats)
    A, n, sx, sxx = _stats_from_q(r["var_x"], obs, p["mask"], starts, T, True, p["prior_tran"],
                                  kind == "niw_full")
    assert_block(s["A"], A, s_rtol, "A")
    assert_block(s["n"], n, s_rtol, "n")
    assert_block(s["sx"], sx, s_rtol, "sx")
    assert_block(s["sxx"], sxx, s_rtol, "sxx")
    assert_block(s["q0"], r["var_x"][:, 0].sum(0), s_rtol, "q0")
    np.testing.assert_allclose(s["logZ"], r["logZ"].sum(), rtol=lz_rtol)
    np.testing.assert_allclose(s["lb_q4"], r["lb"], rtol=lz_rtol)
    assert s["B"] == B
    # the natural-gradient step from those statistics (hmmsgd_metaobs.py:1010-1069)
    Lh, Tf = max(T // 2, 1), obs.shape[0]
    eng.global_update(stats, 0.37, (Tf - 2 * Lh - 1) / (2. * Lh * B), (Tf - 2 * Lh - 1) / ((2. * Lh + 1.) * B))
    vt, vi, em = eng.get_globals()
    e = eng.unpack_emit(em)
    g_rtol = max(s_rtol, S_RTOL)
    assert_block(vt, r["var_tran_new"], g_rtol, "var_tran")
    for key in ("mu", "sigma", "kappa", "nu"):
        ref = np.array([np.broadcast_to(x[key], e[key][0].shape) for x in r["emit_new"]])
        assert_block(e[key], ref, g_rtol, key)
    eng.close()


@pytest.mark.parametrize("onecta", [False, True])
@pytest.mark.parametrize("B", [256, 700, 5000])
def test_c2_bench_shape_float32_series(B, onecta):
    """BASELINE configs[1] as bench.py runs it: float32 series (128-bit vector loads of the window
    rows), B = 256 windows, B = 700 > the 296 CTAs the GPU holds at once and B = 5000 (several waves
    accumulating into the same statistics; the batched path's float32 -> float64 accumulator flushes),
    through the batched tensor-core path (the default) and the one-CTA-per-window pipelined kernel."""
    _run_case(16, 8, 512, B, "niw_diag", "f32", onecta=onecta)


@pytest.mark.parametrize("onecta", [False, True])
def test_c2_bench_shape_float64_series(onecta):
    _run_case(16, 8, 512, 300, "niw_diag", "f64", seed_off=1, onecta=onecta)


def test_c3_bench_shape_float32_series():
    """BASELINE configs[2] window (K64, D32, T1024, full covariance), float32 series, 64 windows."""
    _run_case(64, 32, 1024, 64, "niw_full", "f32", sep=0.3)


@pytest.mark.parametrize("K,D,T,B,kind", [(64, 32, 300, 5, "niw_full"), (32, 16, 200, 7, "niw_diag"),
                                          (20, 4, 129, 3, "niw_full"), (48, 8, 33, 9, "niw_diag"),
                                          (64, 32, 128, 160, "niw_full")])
def test_tensor_core_statistics_path(K, D, T, B, kind):
    """k_stats_tc (tcgen05 + TMA; 16 < K <= 64, float32 series): ragged last tiles (T % 128 != 0), masked
    and NaN rows, more tiles than CTAs (160 windows x 1 tile > 148), full and diagonal second moments."""
    _run_case(K, D, T, B, kind, "f32", sep=0.3, miss=0.1, nan_frac=0.02)


def test_c4_bench_shape_bf16_dense_against_oracle():
    """BASELINE configs[3] as bench.py --config c4 runs it (K256, D64 diagonal, T256, bf16 tensor-core
    recursion), 256 windows = 2 tiles of 128, against the ORACLE: marginals 3e-2 absolute, statistics
    and updated globals 3e-2 relative to the largest entry, log normalisers 1e-2 (restated tolerance
    of the opt-in bf16 path; the reference is float64 only)."""
    from pysvihmm_b200 import _lib as L
    _run_case(256, 64, 256, 256, "niw_diag", "f32", flags_extra=L.BF16_DENSE, sep=0.25, q_abs=3e-2, s_rtol=3e-2,
              lz_rtol=1e-2)


def test_c4_bench_shape_float32_recursions():
    """The same shape without the flag (float32 recursions, generic kernels) meets the 1e-5 bound."""
    _run_case(256, 64, 256, 130, "niw_diag", "f32", sep=0.25, seed_off=3)


@pytest.mark.parametrize("T,B,sep", [(256, 3, 0.8), (256, 3, 0.2), (2048, 2, 0.8)])
def test_c5_bench_shape_gmm(T, B, sep):
    """BASELINE configs[4] (32 states x 4 full-covariance components, D16) at T = 256 with weakly and
    strongly overlapping components (sep 0.8 / 0.2 sigma) and at the benchmark's T = 2048."""
    from oracle import svihmm_oracle as O
    from pysvihmm_b200 import _lib as L
    from pysvihmm_b200.engine import EStepEngine
    K, C, D, kind = 32, 4, 16, "niw_full"
    p = make_random_problem(seed=K * 100 + C + T, K=K * C, D=D, T_full=max(4 * T, 300), kind=kind, miss=0.1, sep=sep)
    rs = np.random.RandomState(K + C)
    emit = [dict(omega=1. + 3. * rs.rand(C), comps=p["emit"][k * C:(k + 1) * C]) for k in range(K)]
    prior = [dict(omega=0.5 + rs.rand(C), comps=p["prior_emit"][k * C:(k + 1) * C]) for k in range(K)]
    var_tran, prior_tran = 1. + 5. * rs.rand(K, K), np.ones((K, K))
    obs = p["obs"].astype(np.float32).astype(np.float64)
    starts = rs.randint(0, obs.shape[0] - T + 1, B)
    eng = EStepEngine(K, D, kind, components=C)
    eng.set_series(obs, p["mask"], dtype="f32")
    eng.set_prior(prior_tran, pack_emit_np(p["prior_emit"]))
    eng.set_mix_weights(np.array([e["omega"] for e in emit]), np.array([e["omega"] for e in prior]))
    eng.set_globals(var_tran, pack_emit_np(p["emit"]))
    vx, stats = eng.estep(starts, T, flags=L.WRAP | L.ADD_PRIOR)
    r = O.gmm_minibatch_step(obs, p["mask"], starts, T, var_tran, emit, prior_tran, prior, 0.37, T // 2,
                             scaled=True)
    assert frac_soft(r["var_x"]) > 0.2, "vacuous parity: posteriors are one-hot"
    assert_q(vx.cpu().numpy(), r["var_x"])
    s = eng.unpack_stats(stats)
    assert_block(s["A"], r["A_inter"], S_RTOL, "A")
    flat = [r["stats"][k][c] for k in range(K) for c in range(C)]
    assert_block(s["n"], np.array([e[1] for e in flat]), S_RTOL, "n")
    assert_block(s["sx"], np.array([e[0] for e in flat]), S_RTOL, "sx")
    assert_block(s["sxx"], np.array([e[2] for e in flat]), S_RTOL, "sxx")
    np.testing.assert_allclose(s["lb_q4"], r["lb"], rtol=3e-6)
    eng.close()


def test_log_domain_tables_match_reference_golden():
    """self.lalpha / self.lbeta (hmmsgd_metaobs.py:775-803, :828-855) rebuilt from the engine's scaled
    tables (svihmm_get_locals + svihmm_get_locals_beta) against the reference's own tables: 1e-6
    relative (+2e-6 absolute: lbeta[T-1] = 0), every entry finite."""
    from pysvihmm_b200 import _lib as L
    from pysvihmm_b200.engine import EStepEngine
    from tests.helpers import SVI_CASES, emit_list, load_golden
    for name in SVI_CASES:
        g = load_golden(name)
        K, D = g["init_tran"].shape[0], g["obs"].shape[1]
        T = 2 * int(g["L"]) + 1
        eng = EStepEngine(K, D)
        eng.set_series(g["obs"], g["mask"], dtype="f64")
        eng.set_globals(g["init_tran"], pack_emit_np(emit_list(g["init_mu"], g["init_sigma"], g["init_kappa"],
                                                               g["init_nu"])))
        starts = g["w_starts"][0]
        eng.estep(starts, T, flags=L.WRAP, keep_locals=True)
        loc = eng.get_locals(len(starts), T)
        for b in range(len(starts)):
            la, lb = eng.log_tables(loc, b)
            assert np.isfinite(la).all() and np.isfinite(lb).all()
            np.testing.assert_allclose(la, g["w_lalpha"][0][b], rtol=1e-6, atol=2e-6)
            np.testing.assert_allclose(lb, g["w_lbeta"][0][b], rtol=1e-6, atol=2e-6)
        eng.close()
    assert Q_RTOL == 1e-5 and Q_ATOL == 2e-7


@pytest.mark.parametrize("K,D,T,B,kind,keep", [(16, 8, 300000, 1, "niw_diag", False), (16, 4, 70001, 1, "niw_full", True),
                                               (5, 3, 6000, 4, "niw_full", True), (2, 2, 1000000, 1, "niw_full", False)])
def test_long_chain_block_parallel_scan(K, D, T, B, kind, keep):
    """full_local_update / batch drivers on ONE long chain (hmmsgd_metaobs.py:1147-1205,
    hmmbase.py:266-320).  K <= 16 and T >= 4096 take the block-parallel scan (scan16.cuh: chunk transfer
    operators on the tensor cores, boundary messages, per-chunk passes) instead of T sequential steps;
    same outputs as the sequential kernels: marginals 1e-5, log normalisers, statistics and - with
    KEEP_LOCALS - the reference's lalpha / lbeta tables at 1e-6 relative."""
    from oracle import svihmm_oracle as O
    from pysvihmm_b200 import _lib as L
    from pysvihmm_b200.engine import EStepEngine
    Tf = B * T + 100
    p = make_random_problem(seed=K * 7 + D, K=K, D=D, T_full=Tf, kind=kind, miss=0.05, sep=0.5)
    starts = np.arange(B) * T + 50
    eng = EStepEngine(K, D, kind)
    eng.set_series(p["obs"], p["mask"], dtype="f64")
    eng.set_prior(p["prior_tran"], pack_emit_np(p["prior_emit"]))
    eng.set_globals(p["var_tran"], pack_emit_np(p["emit"]))
    vx, stats = eng.estep(starts, T, flags=L.MASK_LL | L.ADD_PRIOR, keep_locals=keep)
    r = O.svi_minibatch_step(p["obs"], p["mask"], starts, T, p["var_tran"], p["emit"], p["prior_tran"],
                             p["prior_emit"], 0.5, T // 2, wrap=False, mask_ll=True, scaled=True)
    assert frac_soft(r["var_x"]) > 0.2
    assert_q(vx.cpu().numpy(), r["var_x"])
    s = eng.unpack_stats(stats)
    np.testing.assert_allclose(s["logZ"], r["logZ"].sum(), rtol=3e-6)
    np.testing.assert_allclose(s["lb_q4"], r["lb"], rtol=3e-6)
    A = O.tran_stat(r["var_x"], False).sum(0) + B * (p["prior_tran"] - 1.)
    assert_block(s["A"], A, S_RTOL, "A")
    assert_block(s["n"], np.array([e[1] for e in r["emit_inter"]]), S_RTOL, "n")
    assert_block(s["sx"], np.array([e[0] for e in r["emit_inter"]]), S_RTOL, "sx")
    if keep:
        loc = eng.get_locals(B, T)
        for b in range(B):
            la, lb = eng.log_tables(loc, b)
            assert np.isfinite(la).all() and np.isfinite(lb).all()
            np.testing.assert_allclose(la, r["lalpha"][b], rtol=1e-6, atol=2e-5)
            np.testing.assert_allclose(lb, r["lbeta"][b], rtol=1e-6, atol=2e-5)
    eng.close()


@pytest.mark.parametrize("K,T,B,flags_extra", [(64, 300, 5, 0), (40, 129, 3, 0), (24, 515, 4, 4)])
def test_tensor_core_emissions_edge_cases(K, T, B, flags_extra):
    """k_emit_tc (exact sliced product on tcgen05, D = 32, float32 series): ragged tiles (T not a multiple of
    128), a state count that does not fill the chunks (K = 40: 20 states per warpgroup), NaN rows, masked rows
    with SVIHMM_MASK_LL (flag 4), rows 1e3 times larger / 1e6 times smaller than the rest (per-row scales), an all-zero
    row, a window ending at the last row of the series.  Marginals against the float64 oracle at 1e-5 and
    against the engine's float64 emission kernel (which SVIHMM_KEEP_LOCALS keeps) at the float32 noise floor
    of the recursions."""
    from oracle import svihmm_oracle as O
    from pysvihmm_b200 import _lib as L
    from pysvihmm_b200.engine import EStepEngine
    D, kind = 32, "niw_full"
    p = make_random_problem(seed=K * 1000 + T, K=K, D=D, T_full=4000, kind=kind, miss=0.05, sep=0.4)
    obs = p["obs"].copy()
    starts = np.random.RandomState(5).randint(0, obs.shape[0] - T + 1, B)
    starts[0], starts[-1] = 0, obs.shape[0] - T
    obs[starts[0] + 3] = np.nan
    obs[starts[0] + 17, 5] = np.nan
    obs[starts[1] + 40] *= 1e3           # (beyond ~1e4 the REFERENCE's log-domain tables run out of float64 digits)
    obs[starts[1] + 41] *= 1e-6
    obs[starts[1] + 42] = 0.0
    obs[starts[-1] + T - 1, 0] = np.nan
    obs = obs.astype(np.float32).astype(np.float64)
    mask = p["mask"] | np.isnan(obs).any(1)              # see _run_case: keeps the oracle's sums finite
    eng = EStepEngine(K, D, kind)
    eng.set_series(obs, mask, dtype="f32")
    eng.set_prior(p["prior_tran"], pack_emit_np(p["prior_emit"]))
    eng.set_globals(p["var_tran"], pack_emit_np(p["emit"]))
    fl = L.WRAP | L.ADD_PRIOR | flags_extra
    n0 = eng.launch_count()
    vx, stats = eng.estep(starts, T, flags=fl)
    q = vx.cpu().numpy().copy()
    s = eng.unpack_stats(stats)
    vx64, stats64 = eng.estep(starts, T, flags=fl | L.KEEP_LOCALS)
    q64 = vx64.cpu().numpy()
    assert np.isfinite(q).all()
    assert float(np.max(np.abs(q - q64))) < 2e-6
    if flags_extra == 0:
        r = O.svi_minibatch_step(obs, mask, starts, T, p["var_tran"], p["emit"], p["prior_tran"],
                                 p["prior_emit"], 0.37, max(T // 2, 1), wrap=True, scaled=True)
        assert frac_soft(r["var_x"]) > 0.2, "vacuous parity: posteriors are one-hot"
        assert_q(q, r["var_x"])
        np.testing.assert_allclose(s["logZ"], r["logZ"].sum(), rtol=3e-6)
    s64 = eng.unpack_stats(stats64)
    for key in ("A", "n", "sx", "sxx"):
        assert_block(s[key], s64[key], 1e-5, key)
    eng.close()
