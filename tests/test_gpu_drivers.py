"""The reference-facing classes (same names and call sequence as the reference's own demo
scripts test_hmmsgd_metaobs.py / test_hmmbatchcd.py) against outputs of the reference itself."""
import numpy as np
import pytest

from tests.helpers import SVI_CASES, load_golden

pytestmark = pytest.mark.gpu


def _emit_objs(g, K):
    from pysvihmm_b200.distributions import Gaussian
    return np.array([Gaussian(mu=g["init_mu"][k].copy(), sigma=g["init_sigma"][k].copy(),
                              mu_0=g["prior_mu"], sigma_0=g["prior_sigma"],
                              kappa_0=float(g["prior_kappa"]), nu_0=float(g["prior_nu"]),
                              kappa_mf=float(g["init_kappa"][k]), nu_mf=float(g["init_nu"][k]))
                     for k in range(K)])


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


@pytest.mark.parametrize("name", SVI_CASES)
def test_vbhmm_infer_follows_reference_trajectory(name):
    """hmmsgd_metaobs.VBHMM(...).infer() with the reference's seed: same sampled windows (legacy
    numpy RNG), globals after `maxit` natural-gradient steps within 1e-4 (errors compound over
    steps; per-step parity is 1e-5 in test_gpu_parity)."""
    from pysvihmm_b200 import hmmsgd_metaobs as H
    g = load_golden(name)
    K = g["init_tran"].shape[0]
    seed = {"svi_k3_d2_l5": 11, "svi_k5_d3_l20_mask": 12, "svi_k16_d8_l50": 13, "svi_k2_d2_l1": 14}[name]
    hmm = H.VBHMM(g["obs"].copy(), np.ones(K), np.ones((K, K)), _emit_objs(g, K), tau=1., kappa=0.7,
                  metaobs_half=int(g["L"]), mb_sz=int(g["mb_sz"]), mask=g["mask"],
                  init_tran=g["init_tran"].copy(), maxit=int(g["maxit"]), seed=seed)
    hmm.infer()
    assert _rel(hmm.var_tran, g["g_var_tran"][-1]) < 1e-4
    assert _rel(np.array([e.mu_mf for e in hmm.var_emit]), g["g_mu"][-1]) < 1e-4
    assert _rel(np.array([e.sigma_mf for e in hmm.var_emit]), g["g_sigma"][-1]) < 1e-4
    assert _rel(np.array([e.kappa_mf for e in hmm.var_emit]), g["g_kappa"][-1]) < 1e-4
    assert _rel(np.array([e.nu_mf for e in hmm.var_emit]), g["g_nu"][-1]) < 1e-4
    assert hmm.cur_mo.i1 == g["w_starts"][-1][-1]
    assert hmm.metaobs_fun is None                      # hmmsgd_metaobs.py:485
    # hmmsgd_metaobs.py:273-296 on the device against the host classes' formulas (Dirichlet terms follow
    # the reference line by line; the inverse-Wishart terms are restated on both sides)
    np.testing.assert_allclose(hmm.global_lower_bound(), hmm._global_lower_bound_host(), rtol=1e-10)
    # per-window call path of the reference: local_update(metaobs) + intermediate_pars
    hmm2 = H.VBHMM(g["obs"].copy(), np.ones(K), np.ones((K, K)), _emit_objs(g, K),
                   metaobs_half=int(g["L"]), mb_sz=int(g["mb_sz"]), mask=g["mask"],
                   init_tran=g["init_tran"].copy(), maxit=1, seed=seed)
    s0 = int(g["w_starts"][0][0])
    mo = H.MetaObs(s0, s0 + 2 * int(g["L"]))
    hmm2.local_update(mo)
    assert np.max(np.abs(hmm2.var_x - g["w_var_x"][0][0])) < 1e-6
    np.testing.assert_allclose(hmm2.lliks, g["w_ll"][0][0], rtol=1e-11, atol=1e-11)
    # log-domain tables rebuilt from the engine's scaled ones (float64 sums of float32 scale factors)
    np.testing.assert_allclose(hmm2.lalpha, g["w_lalpha"][0][0], rtol=1e-6, atol=2e-6)
    np.testing.assert_allclose(hmm2.lbeta, g["w_lbeta"][0][0], rtol=1e-6, atol=2e-6)
    assert np.isfinite(hmm2.lbeta).all()
    A_i, e_i = hmm2.intermediate_pars(mo)
    assert _rel(A_i, g["w_A_i"][0][0]) < 1e-5
    assert _rel(np.array([e[0] for e in e_i]), g["w_e1"][0][0]) < 1e-5
    assert _rel(np.array([e[2] for e in e_i]), g["w_e3"][0][0]) < 1e-5


def test_hmmbatchcd_two_cluster_demo():
    """test_hmmbatchcd.py:17-56 of the reference: 500 points N((0,0),I) then 500 N((5,5),I),
    vague NIW prior, flat Dirichlets; with explicit emission inits the Hamming distance is 0."""
    from pysvihmm_b200 import hmmbatchcd as HCD
    from pysvihmm_b200.distributions import Gaussian
    rs = np.random.RandomState(0)
    K, D, T = 2, 2, 1000
    obs = np.vstack([rs.randn(T // 2, D), rs.randn(T // 2, D) + 5.])
    sts = np.repeat([0, 1], T // 2)
    sigma0 = 0.75 * np.cov(obs.T)
    emit = np.array([Gaussian(mu=m, sigma=np.eye(D), mu_0=np.zeros(D), sigma_0=sigma0, kappa_0=0.01,
                              nu_0=4.) for m in (np.array([1., 1.]), np.array([4., 4.]))])
    hmm = HCD.VBHMM(obs, np.ones(K), np.ones((K, K)), emit, maxit=20, sts=sts)
    hmm.infer()
    assert hmm.hamming == 0.0
    np.testing.assert_allclose(hmm.lower_bound(), hmm._lower_bound_host(), rtol=1e-10)     # incl. the initial Dirichlet
    assert len(hmm.elbo_vec) >= 2 and np.all(np.isfinite(hmm.elbo_vec))
    assert np.all(np.diff(hmm.elbo_vec) > -1e-3 * np.abs(hmm.elbo_vec[:-1]))   # CAVI bound does not fall
    assert abs(hmm.var_tran.sum() - (K * K + T - 1)) < 1e-2


def test_batch_cavi_class_matches_reference_golden():
    from pysvihmm_b200 import hmmbatchcd as HCD
    g = load_golden("cavi_k2_d2_t200")
    K = 2
    hmm = HCD.VBHMM(g["obs"].copy(), g["prior_init"], g["prior_tran"], _emit_objs(g, K), mask=g["mask"],
                    maxit=len(g["it_lZ"]), epsilon=0.0)
    hmm.lower_bound = lambda: float(np.random.rand())     # as in make_golden.py: never converge
    hmm.infer()
    assert _rel(hmm.var_tran, g["it_var_tran"][-1]) < 1e-5
    assert _rel(hmm.var_init, g["it_var_init"][-1]) < 1e-5
    assert _rel(np.array([e.sigma_mf for e in hmm.var_emit]), g["it_sigma"][-1]) < 1e-5
    assert np.max(np.abs(hmm.var_x - g["it_var_x"][-1])) < 1e-6


def test_hmmbatchsgd_class_matches_reference_golden():
    """hmmbatchsgd.VBHMM.infer (hmmbatchsgd.py:143-259): batch natural gradient on the device."""
    from pysvihmm_b200 import hmmbatchsgd as HBS
    g = load_golden("bsgd_k3_d2_t150")
    K = 3
    hmm = HBS.VBHMM(g["obs"].copy(), g["prior_init"], g["prior_tran"], _emit_objs(g, K), tau=1., kappa=0.7,
                    mask=g["mask"], init_tran=g["init_var_tran"].copy(), maxit=int(g["maxit"]))
    hmm.lower_bound = lambda: 0.0
    hmm.infer()
    assert _rel(hmm.var_tran, g["it_var_tran"][-1]) < 2e-5
    assert _rel(hmm.var_init, g["it_var_init"][-1]) < 2e-5
    assert _rel(np.array([e.mu_mf for e in hmm.var_emit]), g["it_mu"][-1]) < 2e-5
    assert _rel(np.array([e.sigma_mf for e in hmm.var_emit]), g["it_sigma"][-1]) < 2e-5
    assert _rel(np.array([e.kappa_mf for e in hmm.var_emit]), g["it_kappa"][-1]) < 2e-5
    assert _rel(np.array([e.nu_mf for e in hmm.var_emit]), g["it_nu"][-1]) < 2e-5
    assert np.max(np.abs(hmm.var_x - g["it_var_x"][-1])) < 2e-6


def test_full_local_update_masks_likelihood():
    """hmmsgd_metaobs.py:1147-1205: masked rows carry no evidence."""
    from oracle import svihmm_oracle as O
    from pysvihmm_b200 import hmmsgd_metaobs as H
    g = load_golden("svi_k5_d3_l20_mask")
    K = 5
    hmm = H.VBHMM(g["obs"].copy(), np.ones(K), np.ones((K, K)), _emit_objs(g, K), metaobs_half=20,
                  mb_sz=2, mask=g["mask"], init_tran=g["init_tran"].copy(), maxit=1, seed=1)
    vx = hmm.full_local_update()
    obs = g["obs"].copy()
    obs[g["mask"]] = np.nan
    emit = [dict(mu=g["init_mu"][k], sigma=g["init_sigma"][k], kappa=float(g["init_kappa"][k]),
                 nu=float(g["init_nu"][k])) for k in range(K)]
    r = O.local_update(obs[None], O.stationary_init(g["init_tran"]), g["init_tran"], emit)
    assert np.max(np.abs(vx - r["var_x"][0])) < 2e-6


def test_svihmm_surface_runs():
    """hmmsvi.SVIHMM keeps the reference's method surface (hmmsvi.py:88-204)."""
    from pysvihmm_b200 import hmmsvi
    g = load_golden("svi_k3_d2_l5")
    K = 3
    hmm = hmmsvi.SVIHMM(np.ones(K), g["init_tran"], _emit_objs(g, K), g["obs"].copy())
    T = g["obs"].shape[0]

    def mb_gen():
        for s in (0, 50, 100):
            yield range(s, s + 40)
    hmm.batchfactor = T / 40.
    hmm.infer(mb_gen, maxit=2)
    assert np.all(np.isfinite(hmm.var_tran)) and hmm.var_x.shape == (40, K)
    assert list(next(hmm.allobs_batch())) == list(range(T))
    sts, obs = hmm.generate_obs(10)
    assert sts.shape == (10,) and obs.shape == (10, 2)


def test_vbhmm_categorical_emissions_follow_oracle_trajectory():
    """hmmsgd_metaobs.VBHMM with Categorical emission objects (the reference's Categorical branch,
    hmmsgd_metaobs.py:907-926,1071-1084, does not run as shipped; the oracle restates its intent):
    three natural-gradient steps with the reference's window sampler against the oracle."""
    from oracle import svihmm_oracle as O
    from pysvihmm_b200 import hmmsgd_metaobs as H
    from pysvihmm_b200.distributions import Categorical
    from tests.helpers import make_categorical_problem
    K, C, Lh, S, maxit, seed = 4, 6, 10, 5, 3, 7
    p = make_categorical_problem(seed=3, K=K, C=C, T_full=400, miss=0.1)
    objs = [Categorical(weights=e["alpha"] / e["alpha"].sum(), alphav_0=pe["alpha"].copy(), alpha_mf=e["alpha"].copy())
            for e, pe in zip(p["emit"], p["prior_emit"])]
    hmm = H.VBHMM(p["obs"][:, 0].copy(), np.ones(K), p["prior_tran"], np.array(objs), tau=1., kappa=0.7,
                  metaobs_half=Lh, mb_sz=S, mask=p["mask"], init_tran=p["var_tran"].copy(), maxit=maxit,
                  seed=seed, track_elbo=False)
    hmm.infer()
    # the same windows from the legacy RNG stream (ctor draws, then infer reseeds: hmmsgd_metaobs.py:149,309)
    np.random.seed(seed)
    var_tran, emit = p["var_tran"], p["emit"]
    for it in range(maxit):
        c_vec = np.random.randint(Lh, 400 - 1 - Lh + 1, S)
        r = O.svi_minibatch_step(p["obs"], p["mask"], c_vec - Lh, 2 * Lh + 1, var_tran, emit, p["prior_tran"],
                                 p["prior_emit"], (it + 1.) ** -0.7, Lh, S)
        var_tran, emit = r["var_tran_new"], r["emit_new"]
    assert _rel(hmm.var_tran, var_tran) < 1e-4
    assert _rel(np.array([g._alpha_mf for g in hmm.var_emit]), np.array([e["alpha"] for e in emit])) < 1e-4
    assert abs(sum(hmm.var_emit[0].weights) - 1.) < 1e-12
    # global bound on the device (svihmm_global_bound) against the host classes' formulas
    np.testing.assert_allclose(hmm.global_lower_bound(), hmm._global_lower_bound_host(), rtol=1e-10)


def test_vbhmm_gmm_emissions_follow_oracle_trajectory():
    """hmmsgd_metaobs.VBHMM with MixtureDistribution emission objects (EXTENSION, BASELINE config 5):
    two natural-gradient steps against the oracle."""
    from oracle import svihmm_oracle as O
    from pysvihmm_b200 import hmmsgd_metaobs as H
    from pysvihmm_b200.distributions import Categorical, Gaussian, MixtureDistribution
    from tests.helpers import make_random_problem
    K, C, D, Lh, S, maxit, seed = 3, 2, 2, 8, 4, 2, 5
    p = make_random_problem(seed=4, K=K * C, D=D, T_full=300, kind="niw_full", miss=0.1, sep=0.8)
    rs = np.random.RandomState(1)
    om, om0 = 1. + 3. * rs.rand(K, C), 0.5 + rs.rand(K, C)
    objs = []
    for k in range(K):
        comps = [Gaussian(mu=e["mu"].copy(), sigma=e["sigma"].copy(), mu_0=pe["mu"], sigma_0=pe["sigma"],
                          kappa_0=pe["kappa"], nu_0=pe["nu"], kappa_mf=e["kappa"], nu_mf=e["nu"])
                 for e, pe in zip(p["emit"][k * C:(k + 1) * C], p["prior_emit"][k * C:(k + 1) * C])]
        objs.append(MixtureDistribution(comps, Categorical(weights=np.ones(C) / C, alphav_0=om0[k].copy(),
                                                           alpha_mf=om[k].copy())))
    var_tran = 1. + 5. * rs.rand(K, K)
    hmm = H.VBHMM(p["obs"].copy(), np.ones(K), np.ones((K, K)), np.array(objs, dtype=object), tau=1., kappa=0.7,
                  metaobs_half=Lh, mb_sz=S, mask=p["mask"], init_tran=var_tran.copy(), maxit=maxit, seed=seed,
                  track_elbo=False)
    hmm.infer()
    np.random.seed(seed)
    emit = [dict(omega=om[k], comps=p["emit"][k * C:(k + 1) * C]) for k in range(K)]
    prior = [dict(omega=om0[k], comps=p["prior_emit"][k * C:(k + 1) * C]) for k in range(K)]
    vt = var_tran
    for it in range(maxit):
        c_vec = np.random.randint(Lh, 300 - 1 - Lh + 1, S)
        r = O.gmm_minibatch_step(p["obs"], p["mask"], c_vec - Lh, 2 * Lh + 1, vt, emit, np.ones((K, K)), prior,
                                 (it + 1.) ** -0.7, Lh, S)
        vt, emit = r["var_tran_new"], r["emit_new"]
    assert _rel(hmm.var_tran, vt) < 1e-4
    assert _rel(np.array([m.weights._alpha_mf for m in hmm.var_emit]), np.array([e["omega"] for e in emit])) < 1e-4
    assert _rel(np.array([g.mu_mf for m in hmm.var_emit for g in m.components]),
                np.array([g["mu"] for e in emit for g in e["comps"]])) < 1e-4
    np.testing.assert_allclose(hmm.global_lower_bound(), hmm._global_lower_bound_host(), rtol=1e-10)


def _adaptive_problem():
    from tests.helpers import make_random_problem
    p = make_random_problem(seed=6, K=3, D=2, T_full=400, kind="niw_full", miss=0.0, sep=1.5)
    return p


def _gauss_objs(p):
    from pysvihmm_b200.distributions import Gaussian
    return np.array([Gaussian(mu=e["mu"].copy(), sigma=e["sigma"].copy(), mu_0=pe["mu"], sigma_0=pe["sigma"],
                              kappa_0=pe["kappa"], nu_0=pe["nu"], kappa_mf=e["kappa"], nu_mf=e["nu"])
                     for e, pe in zip(p["emit"], p["prior_emit"])])


def test_select_L_and_select_buffer_match_oracle():
    """select_L / select_buffer (hmmsgd_metaobs.py:521-661) batched over the sampled indices against
    the oracle's per-index loops, same legacy-RNG index draws."""
    from oracle import svihmm_oracle as O
    from pysvihmm_b200 import hmmsgd_metaobs as H
    p = _adaptive_problem()
    K = 3
    hmm = H.VBHMM(p["obs"].copy(), np.ones(K), np.ones((K, K)), _gauss_objs(p), metaobs_half=5, mb_sz=6,
                  init_tran=p["var_tran"].copy(), maxit=1, seed=3)
    var_init = hmm.var_init.copy()                       # prior_init / sum (hmmbase.py:108-111)
    for eps in (1e-2, 1e-4):
        np.random.seed(10)
        L_gpu = hmm.select_L(6, epsilon=eps, minHalfL=1, Lincrement=2, Lcutoff=60)
        np.random.seed(10)
        idx = np.random.choice(400 - 2 * 1 - 1, size=6) + 1
        L_ref = O.select_L(p["obs"], idx, var_init, p["var_tran"], p["emit"], epsilon=eps, minHalfL=1,
                           Lincrement=2, Lcutoff=60)
        assert L_gpu == L_ref, (eps, L_gpu, L_ref)
        np.random.seed(11)
        b_gpu = hmm.select_buffer(5, epsilon=eps, halfL=4, Lincrement=1, Lcutoff=60)
        np.random.seed(11)
        idx = np.random.choice(400 - 2 * 4 - 1, size=5) + 4
        b_ref = O.select_buffer(p["obs"], idx, var_init, p["var_tran"], p["emit"], epsilon=eps, halfL=4,
                                Lincrement=1, Lcutoff=60)
        assert b_gpu == b_ref, (eps, b_gpu, b_ref)
    q = hmm.get_local_messages(100, 7)
    q_ref = O.get_local_messages(p["obs"], 100, 7, var_init, p["var_tran"], p["emit"])
    assert np.max(np.abs(q - q_ref)) < 1e-6
    assert hmm.buffer_budget(10) == 20


def test_buffered_estep_uses_inner_rows_only():
    """svihmm_estep_buffered / intermediate_pars_buffer (hmmsgd_metaobs.py:932-1008)."""
    from oracle import svihmm_oracle as O
    from pysvihmm_b200 import _lib as L
    from pysvihmm_b200.engine import EStepEngine
    from tests.helpers import make_random_problem, pack_emit_np
    for K, D, kind, bufL, Lh in [(3, 2, "niw_full", 9, 4), (40, 3, "niw_full", 12, 5), (5, 4, "niw_diag", 6, 6)]:
        p = make_random_problem(seed=K, K=K, D=D, T_full=300, kind=kind, miss=0.1)
        starts = np.random.RandomState(2).randint(0, 300 - (2 * bufL + 1), 5)
        eng = EStepEngine(K, D, kind)
        eng.set_series(p["obs"], p["mask"], dtype="f64")
        eng.set_prior(p["prior_tran"], pack_emit_np(p["prior_emit"]))
        eng.set_globals(p["var_tran"], pack_emit_np(p["emit"]))
        vx, stats = eng.estep(starts, 2 * bufL + 1, flags=L.WRAP | L.ADD_PRIOR, trim=bufL - Lh)
        if kind == "niw_full":
            r = O.buffered_stats(p["obs"], p["mask"], starts, bufL, Lh, p["var_tran"], p["emit"], p["prior_tran"])
            s = eng.unpack_stats(stats)
            assert np.max(np.abs(vx.cpu().numpy() - r["var_x"])) < 1e-5
            assert _rel(s["A"], r["A_inter"]) < 1e-5
            assert _rel(s["sx"], np.array([e[0] for e in r["emit_inter"]])) < 1e-5
            assert _rel(s["sxx"], np.array([e[2] for e in r["emit_inter"]])) < 1e-5
            assert abs(s["lb_q4"] - r["lb"]) < 3e-6 * abs(r["lb"])
        else:       # trim = 0 must equal the plain call
            vx0, st0 = eng.estep(starts, 2 * bufL + 1, flags=L.WRAP | L.ADD_PRIOR)
            np.testing.assert_allclose(stats.cpu().numpy(), st0.cpu().numpy(), rtol=1e-12)
        eng.close()


def test_vbhmm_infer_adaptive_and_growbuffer_run():
    """infer(adaptive=True) and growBuffer/bufferBudget (hmmsgd_metaobs.py:354-393) end to end."""
    from pysvihmm_b200 import hmmsgd_metaobs as H
    p = _adaptive_problem()
    K = 3
    hmm = H.VBHMM(p["obs"].copy(), np.ones(K), np.ones((K, K)), _gauss_objs(p), metaobs_half=None, mb_sz=4,
                  init_tran=p["var_tran"].copy(), maxit=4, seed=3)
    hmm.infer(adaptive=True, perIter=2, epsilon=1e-3, Lincrement=2, Lcutoff=40)
    assert hmm.metaobs_half >= 1 and np.isfinite(hmm.var_tran).all() and np.isfinite(hmm.elbo_vec).all()
    hmm = H.VBHMM(p["obs"].copy(), np.ones(K), np.ones((K, K)), _gauss_objs(p), metaobs_half=4, mb_sz=4,
                  init_tran=p["var_tran"].copy(), maxit=4, seed=3, growBuffer=True, bufferBudget=True)
    hmm.infer(perIter=2, epsilon=1e-3, Lcutoff=40)
    assert np.isfinite(hmm.var_tran).all() and np.all(hmm.var_tran > 0)
    with pytest.raises(RuntimeError):
        hmm.infer(adaptive=True)


def test_ffbs_samples_follow_the_reference_distribution():
    """svihmm_ffbs / ffbs_fast (hmm_fast.pyx:43-124): the forward table against the oracle's restatement
    (log(A+eps) transition weights, :97-100) and the empirical marginals / pairwise frequencies of 6000
    sampled paths against the exact ones of that distribution (5 sigma binomial bands); different RNG
    than libc rand(), so agreement is in distribution."""
    from oracle import svihmm_oracle as O
    from pysvihmm_b200 import hmmbatchcd as H
    from tests.helpers import make_random_problem
    for K, T in [(3, 40), (40, 25)]:
        p = make_random_problem(seed=K, K=K, D=2, T_full=T, kind="niw_full", miss=0.0, sep=1.0)
        var_init = 0.5 + np.random.RandomState(1).rand(K)
        hmm = H.VBHMM(p["obs"].copy(), np.ones(K), np.ones((K, K)), _gauss_objs(p), init_tran=p["var_tran"].copy())
        lalpha_ref, marg, pair = O.ffbs_tables(p["obs"], var_init, p["var_tran"], p["emit"])
        z, lalpha = hmm.ffbs_fast(var_init, seed=5)
        assert z.shape == (T,) and z.min() >= 0 and z.max() < K
        assert np.max(np.abs(lalpha - lalpha_ref)) < 5e-4 * max(1., np.abs(lalpha_ref).max())
        n = 6000
        zs = hmm._ensure_engine().ffbs(var_init, nsamples=n, seed=11)
        assert zs.shape == (n, T)
        freq = np.stack([(zs == k).mean(0) for k in range(K)], axis=1)            # (T, K)
        band = 5. * np.sqrt(np.maximum(marg * (1 - marg), 1e-4) / n) + 2e-3
        assert np.all(np.abs(freq - marg) < band), float(np.max(np.abs(freq - marg) - band))
        t = T // 2
        pf = np.zeros((K, K))
        np.add.at(pf, (zs[:, t], zs[:, t + 1]), 1. / n)
        bandp = 5. * np.sqrt(np.maximum(pair[t] * (1 - pair[t]), 1e-4) / n) + 2e-3
        assert np.all(np.abs(pf - pair[t]) < bandp)
        assert not np.array_equal(zs[0], zs[1]) or K == 1                         # independent streams


def test_vbhmm_adagrad_follows_reference_trajectory():
    """hmmsgd_metaobs.VBHMM(adagrad=True).infer() against the reference's own run (:1036-1040)."""
    from pysvihmm_b200 import hmmsgd_metaobs as H
    g = load_golden("svi_k3_d2_l5_adagrad")
    K = g["init_tran"].shape[0]
    hmm = H.VBHMM(g["obs"].copy(), np.ones(K), np.ones((K, K)), _emit_objs(g, K), tau=1., kappa=0.7,
                  metaobs_half=int(g["L"]), mb_sz=int(g["mb_sz"]), mask=g["mask"],
                  init_tran=g["init_tran"].copy(), maxit=int(g["maxit"]), seed=15, adagrad=True)
    hmm.infer()
    assert _rel(hmm.var_tran, g["g_var_tran"][-1]) < 1e-4
    assert _rel(np.array([e.mu_mf for e in hmm.var_emit]), g["g_mu"][-1]) < 1e-4
    assert _rel(np.array([e.sigma_mf for e in hmm.var_emit]), g["g_sigma"][-1]) < 1e-4


def test_pred_logprob_matches_oracle():
    """pred_logprob / pred_logprob_full (hmmsgd_metaobs.py:1086-1145)."""
    from oracle import svihmm_oracle as O
    from pysvihmm_b200 import hmmsgd_metaobs as H
    from tests.helpers import make_random_problem
    K = 4
    p = make_random_problem(seed=8, K=K, D=3, T_full=200, kind="niw_full", miss=0.2, sep=1.0)
    hmm = H.VBHMM(p["obs"].copy(), np.ones(K), np.ones((K, K)), _gauss_objs(p), metaobs_half=10, mb_sz=2,
                  mask=p["mask"], init_tran=p["var_tran"].copy(), maxit=1, seed=1)
    mo = H.MetaObs(50, 70)
    got = hmm.pred_logprob(mo)
    r = O.local_update(p["obs"][50:71][None], O.stationary_init(p["var_tran"]), p["var_tran"], p["emit"])
    m = p["mask"][50:71]
    ref = np.mean(np.logaddexp.reduce(np.log(r["var_x"][0][m] + 1e-9) + r["ll"][0][m], axis=1))
    assert abs(got - ref) < 1e-5 * max(1., abs(ref))
    xo = p["obs"].copy(); xo[p["mask"]] = np.nan
    rf = O.local_update(xo[None], O.stationary_init(p["var_tran"]), p["var_tran"], p["emit"])
    ll = O.lliks_gaussian(p["obs"][None], p["emit"])[0]
    mm = p["mask"]
    ref_full = np.mean(np.logaddexp.reduce(np.log(rf["var_x"][0][mm] + 1e-9) + ll[mm], axis=1))
    got_full = hmm.pred_logprob_full()
    assert abs(got_full - ref_full) < 1e-5 * max(1., abs(ref_full))


def test_global_bound_on_device_matches_host_classes():
    """svihmm_global_bound (hmmsgd_metaobs.py:273-296, hmmbase.py:145-199) for the diagonal model and for
    full covariances at D = 32 against the host classes' get_vlb + the Dirichlet terms."""
    from pysvihmm_b200.distributions import DiagonalGaussian, Gaussian
    from pysvihmm_b200.engine import EStepEngine
    from pysvihmm_b200.hmmbase import VariationalHMMBase as Base
    from tests.helpers import make_random_problem, pack_emit_np
    for K, D, kind in [(5, 7, "niw_diag"), (6, 32, "niw_full")]:
        p = make_random_problem(seed=K + D, K=K, D=D, T_full=50, kind=kind)
        prior_init, var_init = 1. + np.random.RandomState(1).rand(K), 0.5 + 3. * np.random.RandomState(2).rand(K)
        eng = EStepEngine(K, D, kind)
        eng.set_prior(p["prior_tran"] + 0.3, pack_emit_np(p["prior_emit"]), prior_init)
        eng.set_globals(p["var_tran"], pack_emit_np(p["emit"]), var_init)
        cls = DiagonalGaussian if kind == "niw_diag" else Gaussian
        vlb = sum(cls(mu=e["mu"], sigma=e["sigma"], mu_0=pe["mu"], sigma_0=pe["sigma"], kappa_0=pe["kappa"],
                      nu_0=pe["nu"], kappa_mf=e["kappa"], nu_mf=e["nu"]).get_vlb()
                  for e, pe in zip(p["emit"], p["prior_emit"]))
        tran = Base._dirichlet_bound(None, p["prior_tran"] + 0.3, p["var_tran"])
        init = Base._dirichlet_bound(None, prior_init, var_init)
        np.testing.assert_allclose(eng.global_bound(False), tran + vlb, rtol=1e-10)
        np.testing.assert_allclose(eng.global_bound(True), tran + init + vlb, rtol=1e-10)
        eng.close()
