#!/usr/bin/env python
"""Generate golden fixtures by running the REFERENCE ITSELF (patched to Python 3
by oracle/build_ref.py; arithmetic untouched) on seeded inputs.

Run in the build container only (needs /root/reference):
    python oracle/build_ref.py && python tests/golden/make_golden.py
The .npz files written next to this script are committed; nothing at test time
reads /root/reference or oracle/_ref.

Every fixture stores the inputs (obs, mask, window starts, initial globals,
priors) and the reference's outputs (lliks, lalpha, lbeta, var_x per window,
A_inter / emission statistics per window, post-update globals).
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")
warnings.filterwarnings("ignore")
sys.path.insert(0, REF)

import hmmsgd_metaobs as HSGD          # noqa: E402
import hmmbatchcd as HCD               # noqa: E402
import hmmbatchsgd as HBS              # noqa: E402
import gen_synthetic as GS             # noqa: E402
from pybasicbayes.distributions import Categorical, Gaussian   # noqa: E402


def make_problem(seed, K, D, T_full, sep, miss=0.0):
    """Synthetic series from reference gen_synthetic.generate_data (gen_synthetic.py:8-56)."""
    rs = np.random.RandomState(seed)
    tran = 0.9 * np.eye(K) + 0.1 / (K - 1) * (1 - np.eye(K))
    mus = sep * rs.randn(K, D)
    emit_true = [Gaussian(mu=mus[k], sigma=np.eye(D), mu_0=np.zeros(D), sigma_0=np.eye(D),
                          kappa_0=1., nu_0=D + 2.) for k in range(K)]
    np.random.seed(seed)
    obs, sts, mask = GS.generate_data(tran, emit_true, T_full, miss=miss)
    if mask is None:
        mask = np.zeros(T_full, bool)
    # initial variational emission parameters: perturbed truth, explicit (no prior sampling)
    init = []
    for k in range(K):
        A = rs.randn(D, D) * 0.3
        init.append(dict(mu=mus[k] + 0.5 * rs.randn(D),
                         sigma=(2.0 * np.eye(D) + A.dot(A.T)),
                         kappa=0.7 + rs.rand(), nu=D + 3. + 2 * rs.rand()))
    prior = dict(mu=np.zeros(D), sigma=0.75 * np.cov(obs.T).reshape(D, D), kappa=0.01, nu=D + 2.)
    init_tran = 1. + 5. * rs.rand(K, K)
    return obs, sts, mask, init, prior, init_tran


def emit_objects(init, prior):
    """Reference objects: mu/sigma explicit, so var_emit = deepcopy has mu_mf=mu,
    sigma_mf=sigma and (kappa_mf, nu_mf) given (distributions.py:195-212)."""
    return np.array([Gaussian(mu=e['mu'].copy(), sigma=e['sigma'].copy(), mu_0=prior['mu'],
                              sigma_0=prior['sigma'], kappa_0=prior['kappa'], nu_0=prior['nu'],
                              kappa_mf=e['kappa'], nu_mf=e['nu']) for e in init])


def pack_emit(prefix, objs, out):
    out[prefix + "_mu"] = np.array([g.mu_mf for g in objs])
    out[prefix + "_sigma"] = np.array([g.sigma_mf for g in objs])
    out[prefix + "_kappa"] = np.array([g.kappa_mf for g in objs], dtype=float)
    out[prefix + "_nu"] = np.array([g.nu_mf for g in objs], dtype=float)


def svi_case(name, seed, K, D, T_full, L, mb_sz, sep, miss=0.0, maxit=2, adagrad=False):
    obs, sts, mask, init, prior, init_tran = make_problem(seed, K, D, T_full, sep, miss)
    prior_emit = emit_objects(init, prior)
    hmm = HSGD.VBHMM(obs.copy(), np.ones(K), np.ones((K, K)), prior_emit, tau=1., kappa=0.7,
                     metaobs_half=L, mb_sz=mb_sz, mask=mask, init_tran=init_tran.copy(),
                     maxit=maxit, seed=seed, adagrad=adagrad)
    out = dict(obs=obs, sts=sts, mask=mask, init_tran=init_tran, prior_tran=np.ones((K, K)),
               prior_mu=prior['mu'], prior_sigma=prior['sigma'], prior_kappa=prior['kappa'],
               prior_nu=prior['nu'], L=L, mb_sz=mb_sz, tau=1., kappa_lr=0.7, maxit=maxit)
    pack_emit("init", hmm.var_emit, out)
    rec = dict(starts=[], ll=[], lalpha=[], lbeta=[], var_x=[], var_init=[], A_i=[],
               e1=[], e2=[], e3=[], lb=[])
    glob = dict(var_tran=[], mu=[], sigma=[], kappa=[], nu=[], lrate=[])

    lu, ip, gu, llb = hmm.local_update, hmm.intermediate_pars, hmm.global_update, hmm.local_lower_bound

    def rec_local_update(metaobs=None):
        lu(metaobs=metaobs)
        rec['starts'].append(metaobs.i1)
        rec['ll'].append(hmm.lliks.copy())
        rec['lalpha'].append(hmm.lalpha.copy())
        rec['lbeta'].append(hmm.lbeta.copy())
        rec['var_x'].append(hmm.var_x.copy())
        rec['var_init'].append(hmm.var_init.copy())

    def rec_ip(metaobs=None):
        A_i, e_i = ip(metaobs)
        rec['A_i'].append(A_i.copy())
        rec['e1'].append(np.array([e[0] for e in e_i]))
        rec['e2'].append(np.array([e[1] for e in e_i], dtype=float))
        rec['e3'].append(np.array([e[2] for e in e_i]))
        return A_i, e_i

    def rec_llb():
        v = llb()
        rec['lb'].append(v)
        return v

    def rec_gu(A_inter, emit_inter):
        gu(A_inter, emit_inter)
        glob['var_tran'].append(hmm.var_tran.copy())
        glob['mu'].append(np.array([g.mu_mf for g in hmm.var_emit]))
        glob['sigma'].append(np.array([g.sigma_mf for g in hmm.var_emit]))
        glob['kappa'].append(np.array([g.kappa_mf for g in hmm.var_emit], dtype=float))
        glob['nu'].append(np.array([g.nu_mf for g in hmm.var_emit], dtype=float))
        glob['lrate'].append(hmm.lrate)

    hmm.local_update, hmm.intermediate_pars = rec_local_update, rec_ip
    hmm.global_update, hmm.local_lower_bound = rec_gu, rec_llb
    # global_lower_bound needs the absent pymattutil (shimmed, unverifiable) -> skip it
    hmm.global_lower_bound = lambda: 0.0
    hmm.infer()
    B = mb_sz
    for k, v in rec.items():
        a = np.array(v)
        out["w_" + k] = a.reshape((maxit, B) + a.shape[1:])
    for k, v in glob.items():
        out["g_" + k] = np.array(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    q = out["w_var_x"]
    print("%-22s K=%d D=%d T=%d B=%d it=%d  frac(max q<0.99)=%.2f  masked=%d" % (
        name, K, D, 2 * L + 1, B, maxit, float(np.mean(q.max(-1) < 0.99)), int(mask.sum())))


def cavi_case(name, seed, T, maxit=3):
    """hmmbatchcd.VBHMM on test_hmmbatchcd.py:17-49-like data (two clusters, here
    1.5 apart so the posteriors are not one-hot), explicit emission inits."""
    K, D = 2, 2
    rs = np.random.RandomState(seed)
    sts = (np.arange(T) >= T // 2).astype(int)
    obs = rs.randn(T, D) + 1.5 * sts[:, None]
    mask = np.zeros(T, bool)
    mask[rs.choice(T, T // 10, replace=False)] = True
    prior = dict(mu=np.zeros(D), sigma=0.75 * np.cov(obs.T), kappa=0.01, nu=4.)
    init = [dict(mu=np.array([-0.3, 0.2]), sigma=np.eye(D) * 2., kappa=0.01, nu=4.),
            dict(mu=np.array([1.0, 1.3]), sigma=np.eye(D) * 2., kappa=0.01, nu=4.)]
    prior_emit = emit_objects(init, prior)
    hmm = HCD.VBHMM(obs.copy(), np.ones(K), np.ones((K, K)), prior_emit, mask=mask, maxit=maxit,
                    epsilon=0.0)
    out = dict(obs=obs, sts=sts, mask=mask, prior_init=np.ones(K), prior_tran=np.ones((K, K)),
               prior_mu=prior['mu'], prior_sigma=prior['sigma'], prior_kappa=prior['kappa'],
               prior_nu=prior['nu'], init_var_init=hmm.var_init.copy(),
               init_var_tran=hmm.var_tran.copy(), maxit=maxit)
    pack_emit("init", hmm.var_emit, out)
    rec = dict(var_x=[], lalpha=[], var_init=[], var_tran=[], mu=[], sigma=[], kappa=[], nu=[], lZ=[])
    gu = hmm.global_update

    def rec_gu():
        gu()
        rec['var_x'].append(hmm.var_x.copy())
        rec['lalpha'].append(hmm.lalpha.copy())
        rec['lZ'].append(np.sum(np.logaddexp.reduce(hmm.lalpha, axis=1)))
        rec['var_init'].append(hmm.var_init.copy())
        rec['var_tran'].append(hmm.var_tran.copy())
        rec['mu'].append(np.array([g.mu_mf for g in hmm.var_emit]))
        rec['sigma'].append(np.array([g.sigma_mf for g in hmm.var_emit]))
        rec['kappa'].append(np.array([g.kappa_mf for g in hmm.var_emit], dtype=float))
        rec['nu'].append(np.array([g.nu_mf for g in hmm.var_emit], dtype=float))

    hmm.global_update = rec_gu
    hmm.lower_bound = lambda: float(len(rec['lZ']))   # get_vlb needs absent pymattutil; never converge
    hmm.infer()
    for k, v in rec.items():
        out["it_" + k] = np.array(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-22s CAVI T=%d iters=%d frac(max q<0.99)=%.2f" % (
        name, T, len(rec['lZ']), float(np.mean(out["it_var_x"].max(-1) < 0.99))))


def bsgd_case(name, seed, T, maxit=3):
    """hmmbatchsgd.VBHMM (batch natural gradient, hmmbatchsgd.py:143-259) on overlapping
    two-cluster data with a mask (rows NaN-ed inside infer, :148-149), explicit emission inits."""
    K, D = 3, 2
    rs = np.random.RandomState(seed)
    sts = (np.arange(T) * K) // T
    obs = rs.randn(T, D) + 1.2 * sts[:, None]
    mask = np.zeros(T, bool)
    mask[rs.choice(T, T // 8, replace=False)] = True
    prior = dict(mu=np.zeros(D), sigma=0.75 * np.cov(obs.T), kappa=0.01, nu=4.)
    init = [dict(mu=np.array([-0.3, 0.2]) + 1.1 * k, sigma=np.eye(D) * (1.5 + 0.2 * k), kappa=0.5 + 0.1 * k,
                 nu=5. + k) for k in range(K)]
    prior_emit = emit_objects(init, prior)
    init_tran = 1. + 3. * rs.rand(K, K)
    hmm = HBS.VBHMM(obs.copy(), np.ones(K), np.ones((K, K)), prior_emit, tau=1., kappa=0.7, mask=mask,
                    init_tran=init_tran.copy(), maxit=maxit)
    out = dict(obs=obs, sts=sts, mask=mask, prior_init=np.ones(K), prior_tran=np.ones((K, K)),
               prior_mu=prior['mu'], prior_sigma=prior['sigma'], prior_kappa=prior['kappa'],
               prior_nu=prior['nu'], init_var_init=hmm.var_init.copy(),
               init_var_tran=hmm.var_tran.copy(), maxit=maxit, tau=1., kappa_lr=0.7)
    pack_emit("init", hmm.var_emit, out)
    rec = dict(var_x=[], var_init=[], var_tran=[], mu=[], sigma=[], kappa=[], nu=[], lrate=[])
    gu = hmm.global_update

    def rec_gu(batch=None):
        gu(batch)
        rec['var_x'].append(hmm.var_x.copy())
        rec['var_init'].append(hmm.var_init.copy())
        rec['var_tran'].append(hmm.var_tran.copy())
        rec['mu'].append(np.array([g.mu_mf for g in hmm.var_emit]))
        rec['sigma'].append(np.array([g.sigma_mf for g in hmm.var_emit]))
        rec['kappa'].append(np.array([g.kappa_mf for g in hmm.var_emit], dtype=float))
        rec['nu'].append(np.array([g.nu_mf for g in hmm.var_emit], dtype=float))
        rec['lrate'].append(hmm.lrate)

    hmm.global_update = rec_gu
    hmm.lower_bound = lambda: 0.0           # get_vlb needs the absent pymattutil
    hmm.pred_logprob = lambda: None
    hmm.infer()
    for k, v in rec.items():
        out["it_" + k] = np.array(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-22s batch SGD T=%d iters=%d frac(max q<0.99)=%.2f" % (
        name, T, len(rec['lrate']), float(np.mean(out["it_var_x"].max(-1) < 0.99))))


def ell_1d_case(name, seed):
    """Gaussian(D=1).expected_log_likelihood values: pins the diagonal-Gaussian
    extension (product of 1-D NIW factors) to distributions.py:351-366."""
    rs = np.random.RandomState(seed)
    x = rs.randn(50, 1) * 2
    mu, sig, kap, nu = rs.randn(6), 0.5 + rs.rand(6) * 3, 0.2 + rs.rand(6), 3. + 4 * rs.rand(6)
    ell = np.empty((6, 50))
    for i in range(6):
        g = Gaussian(mu=mu[i:i + 1], sigma=np.array([[sig[i]]]), mu_0=np.zeros(1),
                     sigma_0=np.eye(1), kappa_0=1., nu_0=3., kappa_mf=kap[i], nu_mf=nu[i])
        ell[i] = g.expected_log_likelihood(x)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), x=x, mu=mu, sigma=sig, kappa=kap,
                        nu=nu, ell=ell)
    print("%-22s 1-D ELL table 6x50" % name)


def cat_case(name, seed):
    """Categorical.expected_log_likelihood (distributions.py:1383-1386) and the global
    natural-gradient step of the Categorical branch (hmmsgd_metaobs.py:1071-1084, run
    literally on reference objects).  The reference's Categorical branch of intermediate_pars
    (:907-926) does not run (its fancy indexing raises), so the statistic itself is unpinned."""
    rs = np.random.RandomState(seed)
    K, C = 5, 7
    alpha = 0.3 + 3 * rs.rand(K, C)
    x = rs.randint(0, C, 80)
    ell = np.empty((K, 80))
    objs = []
    for k in range(K):
        g = Categorical(weights=np.ones(C) / C, alphav_0=np.ones(C) * 0.5, alpha_mf=alpha[k].copy())
        g._alpha_mf = alpha[k].copy()
        ell[k] = g.expected_log_likelihood(x)
        objs.append(g)
    emit_inter = rs.rand(K, C) * 20
    lrate, bfact = 0.37, 2.5
    new = np.empty((K, C))
    for k in range(K):                      # hmmsgd_metaobs.py:1071-1084 verbatim
        G = objs[k]
        nats_old = G._alpha_mf - 1.
        nats_new = (1. - lrate) * nats_old + lrate * bfact * emit_inter[k]
        new[k] = nats_new + 1.
    np.savez_compressed(os.path.join(HERE, name + ".npz"), alpha=alpha, x=x, ell=ell,
                        emit_inter=emit_inter, lrate=lrate, bfact=bfact, alpha_new=new)
    print("%-22s categorical ELL table %dx80" % (name, K))


def gen_case(name, seed):
    """gen_synthetic.generate_data (gen_synthetic.py:8-56) under the legacy global RNG."""
    K, D, T = 4, 3, 400
    rs = np.random.RandomState(seed)
    tran = 0.9 * np.eye(K) + 0.1 / (K - 1) * (1 - np.eye(K))
    mus = rs.randn(K, D)
    A = rs.randn(K, D, D) * 0.3
    sig = np.array([np.eye(D) + a.dot(a.T) for a in A])
    emit = [Gaussian(mu=mus[k], sigma=sig[k], mu_0=np.zeros(D), sigma_0=np.eye(D), kappa_0=1.,
                     nu_0=D + 2.) for k in range(K)]
    np.random.seed(seed)
    obs, sts, mask = GS.generate_data(tran, emit, T, miss=0.1)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), tran=tran, mus=mus, sigmas=sig, T=T,
                        seed=seed, miss=0.1, obs=obs, sts=sts, mask=mask)
    print("%-22s gen_synthetic T=%d masked=%d" % (name, T, int(mask.sum())))


def adaptive_case(name, seed, K, D, T_full, sep):
    """The adaptive-window machinery of hmmsgd_metaobs.VBHMM run by the reference itself:
    get_local_messages (:663-700), select_L (:521-569, non-averaged branch), select_buffer (:579-661)
    and local_update + intermediate_pars_buffer (:932-1008) on buffered windows.  The centre indices
    the reference draws (npr.choice, first RNG call of each function) are re-drawn from the same seed
    and stored, so that the oracle can be driven with them explicitly."""
    obs, sts, mask, init, prior, init_tran = make_problem(seed, K, D, T_full, sep, miss=0.08)
    prior_emit = emit_objects(init, prior)
    hmm = HSGD.VBHMM(obs.copy(), np.ones(K), np.ones((K, K)), prior_emit, tau=1., kappa=0.7,
                     metaobs_half=5, mb_sz=3, mask=mask, init_tran=init_tran.copy(), maxit=1, seed=seed)
    # var_init as infer sets it before every local update (hmmsgd_metaobs.py:413-418)
    A_mean = hmm.var_tran / np.sum(hmm.var_tran, axis=1)[:, None]
    ew, ev = np.linalg.eig(A_mean.T)
    hmm.var_init = np.abs(ev[:, np.argsort(ew)[::-1][0]])
    out = dict(obs=obs, sts=sts, mask=mask, init_tran=init_tran, prior_tran=np.ones((K, K)),
               prior_mu=prior['mu'], prior_sigma=prior['sigma'], prior_kappa=prior['kappa'],
               prior_nu=prior['nu'], var_init=hmm.var_init.copy())
    pack_emit("init", hmm.var_emit, out)
    out["glm_ind"], out["glm_half"] = 100, 7
    out["glm_var_x"] = hmm.get_local_messages(100, 7).copy()
    sl = dict(numIndices=4, epsilon=1e-3, minHalfL=2, Lincrement=2, Lcutoff=40)
    np.random.seed(seed + 1)
    out["selL"] = int(hmm.select_L(**sl))
    np.random.seed(seed + 1)
    out["selL_indices"] = np.random.choice(T_full - 2 * sl["minHalfL"] - 1, size=sl["numIndices"]) + sl["minHalfL"]
    out["selL_args"] = np.array([sl["epsilon"], sl["minHalfL"], sl["Lincrement"], sl["Lcutoff"]])
    sb = dict(numIndices=4, epsilon=1e-3, halfL=5, Lincrement=1, Lcutoff=40)
    np.random.seed(seed + 2)
    out["selB"] = int(hmm.select_buffer(**sb))
    np.random.seed(seed + 2)
    out["selB_indices"] = np.random.choice(T_full - 2 * sb["halfL"] - 1, size=sb["numIndices"]) + sb["halfL"]
    out["selB_args"] = np.array([sb["epsilon"], sb["halfL"], sb["Lincrement"], sb["Lcutoff"]])
    # buffered windows as in infer's growBuffer branch (:371-433)
    L, bufferL = 5, max(int(out["selB"]), 7)
    centres = np.array([40, 133, 250])
    n = 2 * bufferL + 1
    hmm.var_x = np.ones((n, K)) / K
    hmm.lalpha, hmm.lbeta, hmm.lliks = np.empty((n, K)), np.empty((n, K)), np.empty((n, K))
    A_inter = np.zeros((K, K))
    e1, e2, e3, vx, lb = 0., 0., 0., [], 0.
    for c in centres:
        mo = HSGD.MetaObs(int(c - bufferL), int(c + bufferL))
        hmm.cur_mo = mo
        hmm.local_update(metaobs=mo)
        A_i, e_i = hmm.intermediate_pars_buffer(mo, bufferL, L)
        A_inter += A_i
        e1 = e1 + np.array([e[0] for e in e_i])
        e2 = e2 + np.array([e[1] for e in e_i], dtype=float)
        e3 = e3 + np.array([e[2] for e in e_i])
        vx.append(hmm.var_x.copy())
        lb += hmm.local_lower_bound()
    out.update(buf_L=L, buf_bufferL=bufferL, buf_starts=centres - bufferL, buf_var_x=np.array(vx),
               buf_A_inter=A_inter, buf_e1=e1, buf_e2=e2, buf_e3=e3, buf_lb=lb)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-22s K=%d D=%d  select_L=%d select_buffer=%d bufferL=%d" % (name, K, D, out["selL"], out["selB"], bufferL))


def ffbs_case(name, seed, K, D, T, sep, npaths=20000):
    """The reference's native sampler hmm_fast.FFBS (hmm_fast.pyx:43-124, compiled by
    oracle/build_ref.py) on one short series: its forward table (deterministic) and the empirical
    state / pair frequencies of `npaths` sampled paths (libc rand() seeded through srand)."""
    import ctypes
    import hmm_fast
    obs, sts, mask, init, prior, init_tran = make_problem(seed, K, D, T, sep)
    prior_emit = emit_objects(init, prior)
    hmm = HSGD.VBHMM(obs.copy(), np.ones(K), np.ones((K, K)), prior_emit, tau=1., kappa=0.7,
                     metaobs_half=2, mb_sz=1, mask=mask, init_tran=init_tran.copy(), maxit=1, seed=seed)
    A_mean = hmm.var_tran / np.sum(hmm.var_tran, axis=1)[:, None]
    ew, ev = np.linalg.eig(A_mean.T)
    var_init = np.abs(ev[:, np.argsort(ew)[::-1][0]])
    ctypes.CDLL(None).srand(seed)
    z, lalpha = hmm_fast.FFBS(hmm, var_init)
    cnt = np.zeros((T, K))
    pair = np.zeros((T - 1, K, K))
    tt = np.arange(T)
    for _ in range(npaths):
        z, _la = hmm_fast.FFBS(hmm, var_init, lalpha)
        cnt[tt, z] += 1
        pair[tt[:-1], z[:-1], z[1:]] += 1
    out = dict(obs=obs, init_tran=init_tran, var_init=var_init, lalpha=np.array(lalpha), counts=cnt,
               pair_counts=pair, npaths=npaths)
    pack_emit("init", hmm.var_emit, out)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-22s K=%d D=%d T=%d paths=%d  frac(max freq<0.99)=%.2f" % (
        name, K, D, T, npaths, float(np.mean(cnt.max(1) / npaths < 0.99))))


def pred_case(name, seed, K, D, T_full, sep):
    """pred_logprob (hmmsgd_metaobs.py:1086-1119), full_local_update (:1147-1205) and
    pred_logprob_full (:1121-1145) run by the reference.  `obs_full` is never assigned by the reference
    itself (its assignment at :315 is commented out); the caller supplies it, as done here."""
    obs, sts, mask, init, prior, init_tran = make_problem(seed, K, D, T_full, sep, miss=0.2)
    prior_emit = emit_objects(init, prior)
    hmm = HSGD.VBHMM(obs.copy(), np.ones(K), np.ones((K, K)), prior_emit, tau=1., kappa=0.7,
                     metaobs_half=10, mb_sz=2, mask=mask, init_tran=init_tran.copy(), maxit=1, seed=seed)
    A_mean = hmm.var_tran / np.sum(hmm.var_tran, axis=1)[:, None]
    ew, ev = np.linalg.eig(A_mean.T)
    hmm.var_init = np.abs(ev[:, np.argsort(ew)[::-1][0]])
    hmm.obs_full = obs.copy()
    mo = HSGD.MetaObs(50, 70)
    hmm.cur_mo = HSGD.MetaObs(0, 20)
    out = dict(obs=obs, mask=mask, init_tran=init_tran, var_init=hmm.var_init.copy(), mo=np.array([50, 70]))
    pack_emit("init", hmm.var_emit, out)
    out["pred_window"] = hmm.pred_logprob(mo)
    out["full_var_x"] = hmm.full_local_update()
    out["pred_full"] = hmm.pred_logprob_full()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-22s K=%d D=%d T=%d masked=%d  pred_window=%.6f pred_full=%.6f" % (
        name, K, D, T_full, int(mask.sum()), out["pred_window"], out["pred_full"]))


def util_case(name, seed):
    """Host-side helpers of the reference run as shipped: make_mask (util.py:164-192, draws from the
    global numpy RNG), make_mask_prediction (:195-207), munkres_match (:237-277, vendored Munkres)."""
    import util as RU
    from scipy.spatial import distance
    rs = np.random.RandomState(seed)
    K = 5
    sts = rs.choice(K, size=600, p=[0.3, 0.25, 0.2, 0.15, 0.1])
    out = dict(sts=sts)
    np.random.seed(seed)
    out["mask_a"] = RU.make_mask(sts, miss=0.2)
    np.random.seed(seed + 1)
    out["mask_b"] = RU.make_mask(sts, miss=0.1, left=150)
    out["mask_pred"] = RU.make_mask_prediction(sts, miss=0.15)
    perm_true = rs.permutation(K)
    pred = perm_true[sts].copy()
    flip = rs.rand(600) < 0.25
    pred[flip] = rs.choice(K, size=int(flip.sum()))
    match = RU.munkres_match(sts, pred, K)
    out.update(pred=pred, match=match, hamming=distance.hamming(sts, match[pred]))
    # small-K exhaustive matcher, KL between Gaussians, NIW natural -> moment conversion, mvnrand draws
    K2 = 4
    t2 = rs.choice(K2, size=200)
    p2 = rs.permutation(K2)[t2]
    bad = rs.rand(200) < 0.3
    p2[bad] = rs.choice(K2, size=int(bad.sum()))
    out.update(t2=t2, p2=p2, match_seq=RU.match_state_seq(t2, p2, K2))
    A0, A1 = rs.randn(3, 3), rs.randn(3, 3)
    kl_in = dict(mu0=rs.randn(3), sig0=A0.dot(A0.T) + np.eye(3), mu1=rs.randn(3), sig1=A1.dot(A1.T) + 2 * np.eye(3))
    out.update({"kl_" + k: v for k, v in kl_in.items()})
    out["kl"] = RU.KL_gaussian(kl_in["mu0"], kl_in["sig0"], kl_in["mu1"], kl_in["sig1"])
    e = RU.NIW_mf_natural_pars(kl_in["mu0"], kl_in["sig0"], 1.7, 6.5)
    m = RU.NIW_nat2moment_pars(*e)
    out.update(n2m_mu=m[0], n2m_sigma=m[1], n2m_kappa=float(m[2]), n2m_nu=float(m[3]))
    np.random.seed(seed + 2)
    out["mvn"] = RU.mvnrand(kl_in["mu1"], kl_in["sig1"], size=5)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("%-22s masks %d/%d/%d  hamming=%.4f" % (name, out["mask_a"].sum(), out["mask_b"].sum(),
                                                  out["mask_pred"].sum(), out["hamming"]))


if __name__ == "__main__":
    svi_case("svi_k3_d2_l5", seed=11, K=3, D=2, T_full=300, L=5, mb_sz=4, sep=0.6)
    svi_case("svi_k5_d3_l20_mask", seed=12, K=5, D=3, T_full=600, L=20, mb_sz=6, sep=0.5, miss=0.15)
    svi_case("svi_k16_d8_l50", seed=13, K=16, D=8, T_full=1500, L=50, mb_sz=3, sep=0.4, maxit=1)
    svi_case("svi_k2_d2_l1", seed=14, K=2, D=2, T_full=60, L=1, mb_sz=5, sep=0.8, maxit=3)
    svi_case("svi_k3_d2_l5_adagrad", seed=15, K=3, D=2, T_full=300, L=5, mb_sz=4, sep=0.6, maxit=3, adagrad=True)
    cavi_case("cavi_k2_d2_t200", seed=21, T=200)
    bsgd_case("bsgd_k3_d2_t150", seed=22, T=150)
    ell_1d_case("ell_1d", seed=31)
    cat_case("cat_ell", seed=32)
    gen_case("gen_synthetic_k4", seed=8675309)
    adaptive_case("adaptive_k3_d2", seed=41, K=3, D=2, T_full=320, sep=0.6)
    ffbs_case("ffbs_k3_d2_t40", seed=51, K=3, D=2, T=40, sep=0.5)
    pred_case("pred_k4_d3", seed=61, K=4, D=3, T_full=200, sep=0.8)
    util_case("util_helpers", seed=71)
