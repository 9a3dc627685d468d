"""Host-side logic that needs no GPU: samplers, parameter dictionaries, packing layouts,
generator parity with the reference's gen_synthetic, missing-data masks."""
import numpy as np
import pytest

from tests.helpers import load_golden, pack_emit_np


def test_gen_synthetic_bit_identical_to_reference():
    """gen_synthetic.generate_data (gen_synthetic.py:8-56) under the legacy global RNG: the
    series drawn by the package equals the one the reference drew for the same seed."""
    from pysvihmm_b200 import gen_synthetic as GS
    from pysvihmm_b200.distributions import Gaussian
    g = load_golden("gen_synthetic_k4")
    K, D = g["mus"].shape
    emit = [Gaussian(mu=g["mus"][k], sigma=g["sigmas"][k], mu_0=np.zeros(D), sigma_0=np.eye(D),
                     kappa_0=1., nu_0=D + 2.) for k in range(K)]
    np.random.seed(int(g["seed"]))
    obs, sts, mask = GS.generate_data(g["tran"], emit, int(g["T"]), miss=float(g["miss"]))
    assert np.array_equal(sts, g["sts"])
    assert np.array_equal(obs, g["obs"])
    assert np.array_equal(mask, g["mask"])


def test_metaobs_samplers_follow_reference():
    from pysvihmm_b200 import hmmsgd_metaobs as H
    hmm = H.VBHMM.__new__(H.VBHMM)
    np.random.seed(3)
    mb = hmm.metaobs_unif(1000, 7, 50)
    np.random.seed(3)
    c = np.random.randint(7, 1000 - 7, 50)          # hmmsgd_metaobs.py:222
    assert [m.i1 for m in mb] == list(c - 7) and [m.i2 for m in mb] == list(c + 7)
    assert all(0 <= m.i1 and m.i2 <= 999 for m in mb)
    np.random.seed(4)
    mb = hmm.metaobs_noverlap(500, 5, 6)
    assert len(mb) == 7                              # reference returns n+1 windows (:244-253)
    cs = sorted(m.i1 + 5 for m in mb[1:])
    assert np.all(np.diff(cs) > 5)


def test_param_dicts_and_errors():
    from pysvihmm_b200 import hmmsgd_metaobs as H
    from pysvihmm_b200.hmmbase import VariationalHMMBase
    d = H.VBHMM.make_param_dict(1, 2, 3, tau=2., metaobs_half=9)
    assert d["metaobs_half"] == 9 and d["tau"] == 2. and d["mb_sz"] == 1 and d["kappa"] == 0.7
    assert set(VariationalHMMBase.make_param_dict(1, 2, 3)) == {"prior_init", "prior_tran", "prior_emit", "mask"}
    obs = np.zeros((20, 2))
    from pysvihmm_b200.distributions import Gaussian
    pe = [Gaussian(mu=np.zeros(2), sigma=np.eye(2), mu_0=np.zeros(2), sigma_0=np.eye(2), kappa_0=1., nu_0=4.)] * 2
    with pytest.raises(RuntimeError):
        H.VBHMM(obs, np.ones(2), np.ones((2, 2)), pe, metaobs_half=0)
    with pytest.raises(RuntimeError):
        H.VBHMM(obs, np.ones(2), np.ones((2, 2)), pe, metaobs_fun="nope")
    with pytest.raises(RuntimeError):
        H.VBHMM(np.zeros((2, 2, 2)), np.ones(2), np.ones((2, 2)), pe)
    hmm = H.VBHMM(obs, np.ones(2), np.ones((2, 2)), pe, metaobs_half=3, mb_sz=4, seed=1)
    assert hmm.var_x.shape == (7, 2) and hmm.mask.dtype == bool and not hmm.mask.any()
    assert np.allclose(hmm.var_tran, 0.5) and hmm.K == 2 and hmm.T == 20 and hmm.D == 2


def test_masks():
    from pysvihmm_b200 import util
    sts = np.repeat([0, 1, 2], 100)
    np.random.seed(0)
    m = util.make_mask(sts, miss=0.2)
    assert m.sum() == 60 and all(m[sts == k].sum() == 20 for k in range(3))
    mp = util.make_mask_prediction(sts, miss=0.1)
    assert mp[-30:].all() and not mp[:-30].any()
    perm = util.munkres_match(np.array([0, 0, 1, 1, 2, 2]), np.array([2, 2, 0, 0, 1, 1]), 3)
    assert list(perm[np.array([2, 2, 0, 0, 1, 1])]) == [0, 0, 1, 1, 2, 2]


def test_niw_natural_roundtrip_follows_util():
    from pysvihmm_b200 import util
    from pysvihmm_b200.distributions import Gaussian
    rs = np.random.RandomState(0)
    mu, A = rs.randn(3), rs.randn(3, 3)
    sig = A.dot(A.T) + np.eye(3)
    e = util.NIW_mf_natural_pars(mu, sig, 1.7, 6.5)
    assert np.allclose(e[2], sig + 1.7 * np.outer(mu, mu)) and e[3] == 6.5 + 2 + 3   # util.py:28-37
    G = Gaussian(mu=np.zeros(3), sigma=np.eye(3), mu_0=np.zeros(3), sigma_0=np.eye(3), kappa_0=1., nu_0=5.)
    util.NIW_mf_moment_pars(G, *e)
    assert np.allclose(G.mu_mf, mu) and np.allclose(G.sigma_mf, sig) and np.isclose(G.nu_mf, 6.5)
    assert np.allclose(G.sigma, sig / (6.5 - 3 - 1))                                   # util.py:59-60


def test_pack_emit_layout():
    e = [dict(mu=np.arange(2.), sigma=np.array([[2., .5], [.5, 3.]]), kappa=0.3, nu=5.)]
    assert np.array_equal(pack_emit_np(e)[0], [0, 1, 2, .5, .5, 3, .3, 5])
    e = [dict(mu=np.arange(2.), sigma=np.array([2., 3.]), kappa=0.3, nu=np.array([5., 6.]))]
    assert np.array_equal(pack_emit_np(e)[0], [0, 1, 2, 3, .3, .3, 5, 6])


def test_gen_synthetic_variants_and_mmap(tmp_path):
    """generate_data_smoothing / _prediction (gen_synthetic.py:59-155) share the draw order of
    generate_data; the memmap writer/reader round-trips (:158-191)."""
    import numpy as np
    from pysvihmm_b200 import gen_synthetic as GS

    class E(object):
        def __init__(self, mu):
            self.mu = np.asarray(mu, dtype=float)

        def rvs(self, size=None):
            return self.mu + np.random.normal(size=(1, len(self.mu)))

    tran = np.array([[0.9, 0.1], [0.2, 0.8]])
    emit = [E([0., 0.]), E([3., 3.])]
    np.random.seed(3); o1, s1, _ = GS.generate_data(tran, emit, 300)
    np.random.seed(3); o2, s2, m2 = GS.generate_data_smoothing(tran, emit, 300, miss=0.2, left=100)
    np.random.seed(3); o3, s3, m3 = GS.generate_data_prediction(tran, emit, 300, miss=0.1)
    assert np.array_equal(o1, o2) and np.array_equal(o1, o3) and np.array_equal(s1, s3)
    assert m2.dtype == bool and not m2[:100].any() and m2.sum() > 0
    assert m3[-30:].all() and not m3[:-30].any()
    f = str(tmp_path / "obs.dat")
    np.random.seed(4); sts = GS.generate_data_mmap(tran, emit, 250, f, chunk=100)
    mm = GS.read_data_mmap(f, 250, 2)
    assert mm.shape == (250, 2) and sts.shape == (250,)
    chunks = list(GS.read_data_chunks(f, 250, 2, 100))
    assert len(chunks) == 2 and np.array_equal(chunks[1], np.asarray(mm[100:200]))


def test_util_helpers_match_reference_fixture():
    """make_mask / make_mask_prediction bit for bit (same draws from the global numpy RNG) and
    munkres_match to the same Hamming distance as the reference's own util.py (fixture util_helpers,
    made by running it; the reference solves the assignment with its vendored Munkres, here scipy)."""
    from scipy.spatial import distance
    from pysvihmm_b200 import util
    from tests.helpers import load_golden
    g = load_golden("util_helpers")
    sts = g["sts"]
    np.random.seed(71)
    assert np.array_equal(util.make_mask(sts, miss=0.2), g["mask_a"])
    np.random.seed(72)
    assert np.array_equal(util.make_mask(sts, miss=0.1, left=150), g["mask_b"])
    assert np.array_equal(util.make_mask_prediction(sts, miss=0.15), g["mask_pred"])
    match = util.munkres_match(sts, g["pred"], 5)
    assert sorted(match) == list(range(5))
    assert np.isclose(distance.hamming(sts, match[g["pred"]]), float(g["hamming"]))
    assert np.array_equal(match, g["match"])            # unique optimum on this input


def test_util_diagnostics_match_reference_fixture():
    """match_state_seq, KL_gaussian, NIW_nat2moment_pars and mvnrand against the reference's own
    outputs (fixture util_helpers)."""
    from pysvihmm_b200 import util
    from tests.helpers import load_golden
    g = load_golden("util_helpers")
    assert np.array_equal(util.match_state_seq(g["t2"], g["p2"], 4), g["match_seq"])
    kl = util.KL_gaussian(g["kl_mu0"], g["kl_sig0"], g["kl_mu1"], g["kl_sig1"])
    assert np.isclose(kl, float(g["kl"]), rtol=1e-12)
    with pytest.raises(RuntimeError):
        util.KL_gaussian(np.zeros(2), np.eye(2), np.zeros(3), np.eye(3))
    e = util.NIW_mf_natural_pars(g["kl_mu0"], g["kl_sig0"], 1.7, 6.5)
    mu, sigma, kappa, nu = util.NIW_nat2moment_pars(*e)
    assert np.allclose(mu, g["n2m_mu"], rtol=1e-13) and np.allclose(sigma, g["n2m_sigma"], rtol=1e-13)
    assert np.isclose(kappa, float(g["n2m_kappa"])) and np.isclose(nu, float(g["n2m_nu"]))
    np.random.seed(73)
    assert np.allclose(util.mvnrand(g["kl_mu1"], g["kl_sig1"], size=5), g["mvn"], rtol=1e-13, atol=1e-15)


def test_bench_nvlink_counters_degrade_to_a_reason():
    """bench.py reads rank 0's NVLink payload counters around the timed region at N > 1; where NVML (or the
    counters) are missing it must hand back a reason string, never raise (the pool's B200s answer
    NOT_SUPPORTED, profiles/r2_nvlink_counters_unavailable.txt)."""
    import bench
    r = bench.nvlink_kib(0)
    assert isinstance(r, str) or (isinstance(r, list) and len(r) == 2 and all(isinstance(x, int) for x in r))
