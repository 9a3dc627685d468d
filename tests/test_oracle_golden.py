"""The oracle (oracle/svihmm_oracle.py) against outputs of the REFERENCE ITSELF
(fixtures made by tests/golden/make_golden.py from the patched-to-py3 reference).
CPU only.  Tolerances: float64 round-off (1e-10 relative)."""
import numpy as np
import pytest

from oracle import svihmm_oracle as O
from tests.helpers import SVI_CASES, emit_list, golden_prior_emit, load_golden

RT, AT = 1e-10, 1e-12


@pytest.mark.parametrize("name", SVI_CASES)
def test_svi_minibatch_matches_reference(name):
    g = load_golden(name)
    obs, mask = g["obs"], g["mask"]
    L, S = int(g["L"]), int(g["mb_sz"])
    T = 2 * L + 1
    K = g["init_tran"].shape[0]
    var_tran = g["init_tran"].copy()
    emit = emit_list(g["init_mu"], g["init_sigma"], g["init_kappa"], g["init_nu"])
    prior_emit = golden_prior_emit(g, K)
    for it in range(int(g["maxit"])):
        lrate = (it + float(g["tau"])) ** (-float(g["kappa_lr"]))
        assert np.isclose(lrate, g["g_lrate"][it])
        starts = g["w_starts"][it]
        r = O.svi_minibatch_step(obs, mask, starts, T, var_tran, emit, g["prior_tran"],
                                 prior_emit, lrate, L, S)
        np.testing.assert_allclose(r["var_init"], g["w_var_init"][it][0], rtol=1e-9, atol=AT)
        np.testing.assert_allclose(r["ll"], g["w_ll"][it], rtol=RT, atol=AT)
        np.testing.assert_allclose(r["lalpha"], g["w_lalpha"][it], rtol=RT, atol=1e-9)
        np.testing.assert_allclose(r["lbeta"], g["w_lbeta"][it], rtol=RT, atol=1e-9)
        np.testing.assert_allclose(r["var_x"], g["w_var_x"][it], rtol=1e-9, atol=AT)
        # per-window statistics (Q1 product of marginals, Q2 wrap, Q5 prior per window)
        for b in range(S):
            mw = mask[starts[b]:starts[b] + T]
            A_i, e_i = O.intermediate_pars(r["var_x"][b], obs[starts[b]:starts[b] + T], mw,
                                           g["prior_tran"], wrap=True)
            np.testing.assert_allclose(A_i, g["w_A_i"][it][b], rtol=RT, atol=AT)
            np.testing.assert_allclose(np.array([e[0] for e in e_i]), g["w_e1"][it][b], rtol=RT, atol=AT)
            np.testing.assert_allclose(np.array([e[1] for e in e_i]), g["w_e2"][it][b], rtol=RT, atol=AT)
            np.testing.assert_allclose(np.array([e[2] for e in e_i]), g["w_e3"][it][b], rtol=RT, atol=AT)
        np.testing.assert_allclose(r["lb"], np.sum(g["w_lb"][it]), rtol=RT)
        # global natural-gradient step
        np.testing.assert_allclose(r["var_tran_new"], g["g_var_tran"][it], rtol=RT, atol=AT)
        for k in range(K):
            np.testing.assert_allclose(r["emit_new"][k]["mu"], g["g_mu"][it][k], rtol=1e-9, atol=AT)
            np.testing.assert_allclose(r["emit_new"][k]["sigma"], g["g_sigma"][it][k], rtol=1e-9, atol=1e-10)
            np.testing.assert_allclose(r["emit_new"][k]["kappa"], g["g_kappa"][it][k], rtol=RT)
            np.testing.assert_allclose(r["emit_new"][k]["nu"], g["g_nu"][it][k], rtol=RT)
        var_tran, emit = r["var_tran_new"], r["emit_new"]


@pytest.mark.parametrize("name", SVI_CASES)
def test_scaled_form_matches_reference(name):
    """messages_scaled (the matrix-product form used for the BASELINE-size GPU parity cases) against
    the reference's own lalpha / lbeta / var_x tables and against the log-domain restatement."""
    g = load_golden(name)
    L, T = int(g["L"]), 2 * int(g["L"]) + 1
    emit = emit_list(g["init_mu"], g["init_sigma"], g["init_kappa"], g["init_nu"])
    starts = g["w_starts"][0]
    xw = g["obs"][np.asarray(starts)[:, None] + np.arange(T)[None]]
    vi = O.stationary_init(g["init_tran"])
    rs = O.local_update(xw, vi, g["init_tran"], emit, scaled=True)
    np.testing.assert_allclose(rs["lalpha"], g["w_lalpha"][0], rtol=RT, atol=1e-9)
    np.testing.assert_allclose(rs["lbeta"], g["w_lbeta"][0], rtol=RT, atol=1e-9)
    np.testing.assert_allclose(rs["var_x"], g["w_var_x"][0], rtol=1e-9, atol=AT)
    rl = O.local_update(xw, vi, g["init_tran"], emit)
    np.testing.assert_allclose(O.local_lower_bound(rs["lalpha"]), O.local_lower_bound(rl["lalpha"]), rtol=RT)
    np.testing.assert_allclose(O.log_Z(rs["lalpha"]), O.log_Z(rl["lalpha"]), rtol=RT)


def test_wrap_quirk_is_needed():
    """Q2: without the wrap-around term the statistic differs by O(1)."""
    g = load_golden("svi_k3_d2_l5")
    q = g["w_var_x"][0][0]
    A_wrap = g["prior_tran"] + O.tran_stat(q[None], True)[0] - 1.
    A_nowrap = g["prior_tran"] + O.tran_stat(q[None], False)[0] - 1.
    assert np.allclose(A_wrap, g["w_A_i"][0][0], rtol=1e-12)
    assert abs(A_wrap.sum() - A_nowrap.sum() - 1.0) < 1e-9


def test_batch_cavi_matches_reference():
    g = load_golden("cavi_k2_d2_t200")
    K = 2
    var_init, var_tran = g["init_var_init"], g["init_var_tran"]
    emit = emit_list(g["init_mu"], g["init_sigma"], g["init_kappa"], g["init_nu"])
    prior_emit = golden_prior_emit(g, K)
    for it in range(len(g["it_lZ"])):
        r = O.batch_cavi_step(g["obs"], g["mask"], var_init, var_tran, emit, g["prior_init"],
                              g["prior_tran"], prior_emit)
        np.testing.assert_allclose(r["var_x"][0], g["it_var_x"][it], rtol=1e-9, atol=AT)
        np.testing.assert_allclose(r["lZ"], g["it_lZ"][it], rtol=RT)
        np.testing.assert_allclose(r["var_init_new"], g["it_var_init"][it], rtol=RT)
        np.testing.assert_allclose(r["var_tran_new"], g["it_var_tran"][it], rtol=RT)
        for k in range(K):
            np.testing.assert_allclose(r["emit_new"][k]["mu"], g["it_mu"][it][k], rtol=1e-9, atol=AT)
            np.testing.assert_allclose(r["emit_new"][k]["sigma"], g["it_sigma"][it][k], rtol=1e-9, atol=AT)
            np.testing.assert_allclose(r["emit_new"][k]["kappa"], g["it_kappa"][it][k], rtol=RT)
            np.testing.assert_allclose(r["emit_new"][k]["nu"], g["it_nu"][it][k], rtol=RT)
        var_init, var_tran, emit = r["var_init_new"], r["var_tran_new"], r["emit_new"]


def test_batch_sgd_matches_reference():
    """hmmbatchsgd.VBHMM.infer iterations (hmmbatchsgd.py:143-259)."""
    g = load_golden("bsgd_k3_d2_t150")
    K = 3
    var_init, var_tran = g["init_var_init"], g["init_var_tran"]
    emit = emit_list(g["init_mu"], g["init_sigma"], g["init_kappa"], g["init_nu"])
    prior_emit = golden_prior_emit(g, K)
    for it in range(int(g["maxit"])):
        lrate = (it + float(g["tau"])) ** (-float(g["kappa_lr"]))
        assert np.isclose(lrate, g["it_lrate"][it])
        r = O.batch_sgd_step(g["obs"], g["mask"], var_init, var_tran, emit, g["prior_init"],
                             g["prior_tran"], prior_emit, lrate)
        np.testing.assert_allclose(r["var_x"][0], g["it_var_x"][it], rtol=1e-9, atol=AT)
        np.testing.assert_allclose(r["var_init_new"], g["it_var_init"][it], rtol=RT)
        np.testing.assert_allclose(r["var_tran_new"], g["it_var_tran"][it], rtol=RT)
        for k in range(K):
            np.testing.assert_allclose(r["emit_new"][k]["mu"], g["it_mu"][it][k], rtol=1e-9, atol=AT)
            np.testing.assert_allclose(r["emit_new"][k]["sigma"], g["it_sigma"][it][k], rtol=1e-9, atol=1e-10)
            np.testing.assert_allclose(r["emit_new"][k]["kappa"], g["it_kappa"][it][k], rtol=RT)
            np.testing.assert_allclose(r["emit_new"][k]["nu"], g["it_nu"][it][k], rtol=RT)
        var_init, var_tran, emit = r["var_init_new"], r["var_tran_new"], r["emit_new"]


def test_diag_extension_pinned_to_1d_reference():
    g = load_golden("ell_1d")
    for i in range(6):
        got = O.diag_gaussian_ell(g["x"], g["mu"][i:i + 1], g["sigma"][i:i + 1],
                                  g["kappa"][i:i + 1], g["nu"][i:i + 1])
        np.testing.assert_allclose(got, g["ell"][i], rtol=1e-12, atol=1e-12)


def test_logZ_identities():
    """Property tests the reference implies: rows of var_x sum to 1; the Q4 bound is the
    prefix sum of per-step normalisers; logZ = logsumexp lalpha[T-1]."""
    g = load_golden("svi_k5_d3_l20_mask")
    la = g["w_lalpha"][0]
    assert np.allclose(g["w_var_x"][0].sum(-1), 1.0, atol=1e-12)
    assert np.allclose(O.local_lower_bound(la), g["w_lb"][0], rtol=1e-12)
    assert np.all(O.log_Z(la) < 0)


def test_categorical_ell_and_update_match_reference():
    """Categorical.expected_log_likelihood (distributions.py:1383-1386) and the Categorical branch of
    global_update (hmmsgd_metaobs.py:1071-1084) as run on reference objects."""
    g = load_golden("cat_ell")
    K = g["alpha"].shape[0]
    emit = [dict(alpha=g["alpha"][k]) for k in range(K)]
    ll = O.lliks_categorical(g["x"][None].astype(float), emit)[0]
    np.testing.assert_allclose(ll.T, g["ell"], rtol=RT, atol=AT)
    xs = g["x"].astype(float)[None].copy()
    xs[0, 3] = np.nan
    assert np.all(O.lliks_categorical(xs, emit)[0, 3] == 0.)
    # emit_inter = n_windows*(prior-1) + counts: feed the fixture's emit_inter through counts with
    # a prior of ones (prior - 1 = 0)
    for k in range(K):
        new = O.cat_global_update(g["alpha"][k], np.ones_like(g["alpha"][k]), g["emit_inter"][k], 3,
                                  float(g["lrate"]), float(g["bfact"]))
        np.testing.assert_allclose(new, g["alpha_new"][k], rtol=RT, atol=AT)


def test_adagrad_branch_matches_reference():
    """hmmsgd_metaobs.py:1036-1040 (adagrad=True): transition step scaled by ada_G**0.25."""
    g = load_golden("svi_k3_d2_l5_adagrad")
    obs, mask = g["obs"], g["mask"]
    L, S = int(g["L"]), int(g["mb_sz"])
    K = g["init_tran"].shape[0]
    var_tran = g["init_tran"].copy()
    emit = emit_list(g["init_mu"], g["init_sigma"], g["init_kappa"], g["init_nu"])
    prior_emit = golden_prior_emit(g, K)
    ada_G = np.ones((K, K))                                   # :183
    for it in range(int(g["maxit"])):
        r = O.svi_minibatch_step(obs, mask, g["w_starts"][it], 2 * L + 1, var_tran, emit, g["prior_tran"],
                                 prior_emit, float(g["g_lrate"][it]), L, S, ada_G=ada_G)
        np.testing.assert_allclose(r["var_tran_new"], g["g_var_tran"][it], rtol=RT, atol=AT)
        np.testing.assert_allclose(np.array([e["mu"] for e in r["emit_new"]]), g["g_mu"][it], rtol=1e-9, atol=AT)
        var_tran, emit = r["var_tran_new"], r["emit_new"]


def test_adaptive_window_machinery_matches_reference():
    """get_local_messages / select_L / select_buffer / intermediate_pars_buffer of the reference
    (hmmsgd_metaobs.py:521-700,932-1008; fixture adaptive_k3_d2 made by running them) against the
    oracle's restatements, driven with the centre indices the reference drew."""
    g = load_golden("adaptive_k3_d2")
    obs, mask = g["obs"], g["mask"]
    K = g["init_tran"].shape[0]
    var_tran, var_init = g["init_tran"], g["var_init"]
    emit = emit_list(g["init_mu"], g["init_sigma"], g["init_kappa"], g["init_nu"])
    np.testing.assert_allclose(O.stationary_init(var_tran), var_init, rtol=1e-9, atol=AT)
    vx = O.get_local_messages(obs, int(g["glm_ind"]), int(g["glm_half"]), var_init, var_tran, emit)
    np.testing.assert_allclose(vx, g["glm_var_x"], rtol=1e-9, atol=AT)
    eps, minL, inc, cut = g["selL_args"]
    assert O.select_L(obs, g["selL_indices"], var_init, var_tran, emit, epsilon=float(eps), minHalfL=int(minL),
                      Lincrement=int(inc), Lcutoff=int(cut)) == int(g["selL"])
    eps, halfL, inc, cut = g["selB_args"]
    assert O.select_buffer(obs, g["selB_indices"], var_init, var_tran, emit, epsilon=float(eps), halfL=int(halfL),
                           Lincrement=int(inc), Lcutoff=int(cut)) == int(g["selB"])
    r = O.buffered_stats(obs, mask, g["buf_starts"], int(g["buf_bufferL"]), int(g["buf_L"]), var_tran, emit,
                         g["prior_tran"])
    np.testing.assert_allclose(r["var_x"], g["buf_var_x"], rtol=1e-9, atol=AT)
    np.testing.assert_allclose(r["A_inter"], g["buf_A_inter"], rtol=RT, atol=AT)
    np.testing.assert_allclose(np.array([e[0] for e in r["emit_inter"]]), g["buf_e1"], rtol=RT, atol=AT)
    np.testing.assert_allclose(np.array([e[1] for e in r["emit_inter"]]), g["buf_e2"], rtol=RT, atol=AT)
    np.testing.assert_allclose(np.array([e[2] for e in r["emit_inter"]]), g["buf_e3"], rtol=RT, atol=AT)
    np.testing.assert_allclose(r["lb"], float(g["buf_lb"]), rtol=RT)


def test_ffbs_tables_match_reference_sampler():
    """oracle.ffbs_tables against the reference's own native sampler (hmm_fast.FFBS, compiled from
    hmm_fast.pyx by oracle/build_ref.py; fixture ffbs_k3_d2_t40): the forward table it returns
    (log(A + eps) transition weights, hmm_fast.pyx:82-100) to round-off, and the state / pair
    frequencies of its 20000 sampled paths within 5-sigma binomial bands of the exact marginals of
    the law the oracle says it samples from."""
    g = load_golden("ffbs_k3_d2_t40")
    emit = emit_list(g["init_mu"], g["init_sigma"], g["init_kappa"], g["init_nu"])
    lalpha, marg, pair = O.ffbs_tables(g["obs"], g["var_init"], g["init_tran"], emit)
    np.testing.assert_allclose(lalpha, g["lalpha"], rtol=RT, atol=1e-9)
    n = float(g["npaths"])
    for p, cnt in ((marg, g["counts"]), (pair, g["pair_counts"])):
        sd = np.sqrt(np.maximum(p * (1. - p), 1e-12) / n)
        assert np.all(np.abs(cnt / n - p) <= 5. * sd + 2. / n), float(np.max(np.abs(cnt / n - p) / sd))
    assert np.mean(marg.max(1) < 0.99) > 0.5            # not a vacuous (one-hot) case


def test_pred_logprob_matches_reference():
    """pred_logprob / full_local_update / pred_logprob_full of the reference (hmmsgd_metaobs.py:
    1086-1205; fixture pred_k4_d3) against the oracle's restatements."""
    g = load_golden("pred_k4_d3")
    obs, mask, var_tran, var_init = g["obs"], g["mask"], g["init_tran"], g["var_init"]
    emit = emit_list(g["init_mu"], g["init_sigma"], g["init_kappa"], g["init_nu"])
    i1, i2 = (int(v) for v in g["mo"])
    r = O.local_update(obs[i1:i2 + 1][None], var_init, var_tran, emit)    # infer does not NaN the masked rows
    got = O.pred_logprob(r["var_x"][0], obs[i1:i2 + 1], mask[i1:i2 + 1], emit)
    np.testing.assert_allclose(got, float(g["pred_window"]), rtol=RT)
    vx = O.full_local_update(obs, mask, var_init, var_tran, emit)
    np.testing.assert_allclose(vx, g["full_var_x"], rtol=1e-9, atol=AT)
    np.testing.assert_allclose(O.pred_logprob(vx, obs, mask, emit), float(g["pred_full"]), rtol=RT)
    assert O.pred_logprob(vx, obs, np.zeros_like(mask), emit) is None
