"""Float64 CPU oracle for the SVI-HMM local E-step path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module, and only as the
checker.  The product package (pysvihmm_b200) never imports it.

It restates, in plain numpy float64 and in the reference's own *log domain*
(np.logaddexp.reduce, no scaling table), the algorithm of dillonalaird/pysvihmm
for the path named by BASELINE.json.  Every function cites the reference
file:line it follows (paths relative to /root/reference).  The recursions loop
over t in Python exactly like the reference but are vectorised over a leading
batch axis of windows so that B windows cost one pass.

PINNING: the reference ships no golden vectors, so this
oracle is pinned against outputs of the reference itself: oracle/build_ref.py
makes the reference importable on Python 3 (mechanical patches only),
tests/golden/make_golden.py runs reference classes hmmsgd_metaobs.VBHMM /
hmmbatchcd.VBHMM on seeded inputs and commits their outputs as fixtures, and
tests/test_oracle_golden.py checks this file against them (agreement ~1e-13).
Extensions that have no reference implementation (diagonal Gaussian, GMM
emissions) are pinned only indirectly through Gaussian(D=1) calls and say
"parity unpinned" where that is so.
"""
import numpy as np
from scipy.special import digamma

EPS = 1e-9            # hmmbase.py:30, hmmsgd_metaobs.py:26
WEPS = 1e-12          # pybasicbayes/distributions.py:22


# --------------------------------------------------------------------------
# globals -> per-step constants
# --------------------------------------------------------------------------
def stationary_init(var_tran):
    """hmmsgd_metaobs.py:413-418.  |top eigenvector of A_mean^T|, unit L2 norm
    (quirk Q3: it is *not* renormalised to a distribution)."""
    A_mean = var_tran / np.sum(var_tran, axis=1)[:, None]
    ew, ev = np.linalg.eig(A_mean.T)
    ew_dec = np.argsort(ew)[::-1]
    return np.abs(ev[:, ew_dec[0]])


def mod_params(var_init, var_tran):
    """hmmsgd_metaobs.py:502-504 / hmmbase.py:214-216."""
    mod_init = digamma(var_init + EPS) - digamma(np.sum(var_init) + EPS)
    tran_sum = np.sum(var_tran, axis=1)
    mod_tran = digamma(var_tran + EPS) - digamma(tran_sum[:, None] + EPS)
    return mod_init, mod_tran


# --------------------------------------------------------------------------
# emission expected log-likelihoods
# --------------------------------------------------------------------------
def gaussian_loglmbdatilde(sigma_mf, nu_mf):
    """pybasicbayes/distributions.py:361-366."""
    D = sigma_mf.shape[0]
    chol = np.linalg.cholesky(sigma_mf)
    return digamma((nu_mf - np.arange(D)) / 2.).sum() + D * np.log(2) \
        - 2 * np.log(chol.diagonal()).sum()


def gaussian_ell(x, mu_mf, sigma_mf, kappa_mf, nu_mf):
    """pybasicbayes/distributions.py:351-359 (Gaussian.expected_log_likelihood).
    x: (..., D) -> (...)."""
    D = len(mu_mf)
    shp = x.shape[:-1]
    xc = np.reshape(x, (-1, D)) - mu_mf
    xs = np.linalg.solve(np.linalg.cholesky(sigma_mf), xc.T)
    out = gaussian_loglmbdatilde(sigma_mf, nu_mf) / 2 - D / (2 * kappa_mf) \
        - nu_mf / 2 * np.einsum('ij,ij->j', xs, xs) - D / 2 * np.log(2 * np.pi)
    return out.reshape(shp)


def diag_gaussian_ell(x, mu_mf, sig_mf, kappa_mf, nu_mf):
    """EXTENSION (BASELINE config 2; no mean-field DiagonalGaussian exists in the
    reference, pybasicbayes/distributions.py:666-797 is Gibbs/ML only).  Defined
    as a product of D independent one-dimensional NIW factors, i.e.
    sum_d Gaussian_{1-D}.expected_log_likelihood(x_d) with formula :351-366 at
    D=1.  mu_mf, sig_mf, kappa_mf, nu_mf: (D,) each.  Pinned only through
    reference Gaussian(D=1) calls; end-to-end parity unpinned."""
    out = 0.
    for d in range(x.shape[-1]):
        out = out + gaussian_ell(x[..., d:d + 1], mu_mf[d:d + 1],
                                 np.array([[sig_mf[d]]]), kappa_mf[d], nu_mf[d])
    return out


def categorical_ell(x, alpha_mf):
    """pybasicbayes/distributions.py:1383-1386."""
    return digamma(alpha_mf[x]) - digamma(alpha_mf.sum())


def lliks_categorical(xw, emit):
    """hmmsgd_metaobs.py:508-509 with Categorical emissions (distributions.py:1383-1386).
    xw: (B,T) symbols as floats (NaN = missing -> ll = 0, np.nan_to_num semantics);
    emit: list of K dicts(alpha=(C,)) -> (B,T,K)."""
    bad = np.isnan(xw)
    sym = np.where(bad, 0, xw).astype(int)
    ll = np.stack([categorical_ell(sym, e['alpha']) for e in emit], axis=-1)
    ll[bad] = 0.
    return ll


def cat_suffstats(xw, w, C):
    """Intended statistic of the (non-running) Categorical branch hmmsgd_metaobs.py:907-926:
    weighted counts sum_t w_t 1[x_t = c]."""
    out = np.zeros(C)
    np.add.at(out, xw.astype(int), w)
    return out


def cat_global_update(alpha, alpha_prior, counts, n_windows, lrate, bfact):
    """hmmsgd_metaobs.py:1071-1084 with emit_inter[k] = sum over the minibatch of
    (alphav_0 + counts_window - 1) (:925-926): the prior enters once per window and is scaled
    by bfact like the data (quirk Q5 for emissions)."""
    emit_inter = n_windows * (alpha_prior - 1.) + counts
    return (1. - lrate) * (alpha - 1.) + lrate * bfact * emit_inter + 1.


def lliks_gmm(xw, emit):
    """EXTENSION (BASELINE config 5; the reference's MixtureDistribution, pybasicbayes/models.py:
    256-300, has no mean-field path).  State k emits from C NIW components with Dirichlet(omega_k)
    weights.  Following the variational mixture update pybasicbayes/internals/labels.py:52-65
    (logr = E[ln pi] + component expected log-likelihoods, r = softmax):
        ll[t,k]  = logsumexp_c( psi(omega_kc) - psi(sum_c omega_kc) + ELL_kc(x_t) )
        r[t,k,c] = exp(that term - ll[t,k])
    with distributions.py:351-366 per component and :1383-1386 for the weights; NaN rows give
    ll = 0 (hmmsgd_metaobs.py:508-509).  emit: list of K dicts(omega=(C,), comps=[C dicts]).
    Returns ll (B,T,K), r (B,T,K,C).  End-to-end parity unpinned (C = 1 reduces to the pinned
    Gaussian path)."""
    B, T, D = xw.shape
    K, C = len(emit), len(emit[0]['comps'])
    bad = np.isnan(xw).any(-1)
    ll = np.empty((B, T, K)); r = np.empty((B, T, K, C))
    with np.errstate(invalid='ignore'):
        for k, e in enumerate(emit):
            lw = digamma(e['omega']) - digamma(np.sum(e['omega']))
            sc = np.stack([lw[c] + (gaussian_ell(xw, g['mu'], g['sigma'], g['kappa'], g['nu'])
                                    if np.ndim(g['sigma']) == 2 else
                                    diag_gaussian_ell(xw, g['mu'], g['sigma'], g['kappa'], g['nu']))
                           for c, g in enumerate(e['comps'])], axis=-1)
            m = np.max(sc, axis=-1, keepdims=True)
            ll[:, :, k] = (m + np.log(np.sum(np.exp(sc - m), axis=-1, keepdims=True)))[..., 0]
            r[:, :, k] = np.exp(sc - ll[:, :, k][..., None])
    ll[bad] = 0.
    r[bad] = 0.
    return ll, r


def gmm_minibatch_step(obs, mask, starts, T, var_tran, emit, prior_tran, prior_emit, lrate, L, S=None,
                       wrap=True, scaled=False):
    """svi_minibatch_step with GMM emissions (EXTENSION, see lliks_gmm).  Component (k,c) collects
    the NIW statistics (util.py:73-83) weighted by q[t,k] r[t,k,c]; its natural-gradient step is
    hmmsgd_metaobs.py:1048-1069 verbatim; the Dirichlet weights move like the transition rows
    (:1029-1045): omega <- (1-rho)(omega-1) + rho (omega0 - 1 + bE n_kc) + 1."""
    T_full = obs.shape[0]
    K, C = len(emit), len(emit[0]['comps'])
    S = len(starts) if S is None else S
    idx = np.asarray(starts)[:, None] + np.arange(T)[None]
    xw = obs[idx]
    mw = mask[idx] if mask is not None else np.zeros(idx.shape, bool)
    var_init = stationary_init(var_tran)
    mod_init, mod_tran = mod_params(var_init, var_tran)
    ll, r = lliks_gmm(xw, emit)
    if scaled:
        lalpha, lbeta = messages_scaled(ll, mod_init, mod_tran)
    else:
        lalpha = forward_msgs(ll, mod_init, mod_tran)
        lbeta = backward_msgs(ll, mod_tran)
    q = marginals(lalpha, lbeta)
    A_inter = np.zeros_like(var_tran)
    full = np.ndim(emit[0]['comps'][0]['sigma']) == 2
    stats = [[None] * C for _ in range(K)]
    for b in range(len(starts)):
        A_inter += prior_tran + tran_stat(q[b][None], wrap)[0] - 1.
        inds = np.logical_not(mw[b]) & ~np.isnan(xw[b]).any(-1)
        xb = xw[b][inds]
        for k in range(K):
            for c in range(C):
                w = q[b][inds, k] * r[b][inds, k, c]
                e = niw_suffstats(xb, w) if full else diag_suffstats(xb, w)
                stats[k][c] = e if stats[k][c] is None else [u + v for u, v in zip(stats[k][c], e)]
    bA = (T_full - 2 * L - 1) / (2. * L * S)
    bE = (T_full - 2 * L - 1) / ((2. * L + 1.) * S)
    var_tran_new = (1. - lrate) * (var_tran - 1.) + lrate * bA * A_inter + 1.
    nat, mom = (niw_natural, niw_moment) if full else (diag_natural, diag_moment)
    emit_new = []
    for k in range(K):
        comps, om = [], np.empty(C)
        for c in range(C):
            g, p = emit[k]['comps'][c], prior_emit[k]['comps'][c]
            old = nat(g['mu'], g['sigma'], g['kappa'], g['nu'])
            pri = nat(p['mu'], p['sigma'], p['kappa'], p['nu'])
            comps.append(mom(*[(1. - lrate) * o + lrate * (pp + bE * e)
                               for o, pp, e in zip(old, pri, stats[k][c])]))
            om[c] = (1. - lrate) * (emit[k]['omega'][c] - 1.) + lrate * (
                prior_emit[k]['omega'][c] - 1. + bE * stats[k][c][1]) + 1.
        emit_new.append(dict(omega=om, comps=comps))
    return dict(ll=ll, resp=r, lalpha=lalpha, lbeta=lbeta, var_x=q, A_inter=A_inter, stats=stats,
                lb=float(np.sum(local_lower_bound(lalpha))), logZ=log_Z(lalpha),
                var_tran_new=var_tran_new, emit_new=emit_new, var_init=var_init)


def lliks_gaussian(xw, emit):
    """hmmsgd_metaobs.py:508-509: per state expected_log_likelihood, then
    np.nan_to_num (a NaN row gives ll = 0 = 'missing').
    xw: (B,T,D); emit: list of K dicts(mu, sigma, kappa, nu) -> (B,T,K)."""
    B, T, D = xw.shape
    ll = np.empty((B, T, len(emit)))
    with np.errstate(invalid='ignore'):
        for k, e in enumerate(emit):
            if np.ndim(e['sigma']) == 2:
                ll[:, :, k] = gaussian_ell(xw, e['mu'], e['sigma'], e['kappa'], e['nu'])
            else:
                ll[:, :, k] = diag_gaussian_ell(xw, e['mu'], e['sigma'], e['kappa'], e['nu'])
    return np.nan_to_num(ll)


# --------------------------------------------------------------------------
# messages (log domain, exactly the reference's recursions)
# --------------------------------------------------------------------------
def forward_msgs(ll, mod_init, mod_tran):
    """hmmsgd_metaobs.py:775-803 / hmmbase.py:266-295.  ll: (B,T,K)."""
    B, T, K = ll.shape
    lalpha = np.empty((B, T, K))
    lalpha[:, 0] = mod_init + ll[:, 0]
    ltT = mod_tran.T
    for t in range(1, T):
        # lalpha[t,j] = logsumexp_i(lalpha[t-1,i] + ltran[i,j]) + ll[t,j]
        lalpha[:, t] = np.logaddexp.reduce(lalpha[:, t - 1][:, None, :] + ltT[None], axis=2) + ll[:, t]
    return lalpha


def backward_msgs(ll, mod_tran):
    """hmmsgd_metaobs.py:828-855 / hmmbase.py:297-320."""
    B, T, K = ll.shape
    lbeta = np.empty((B, T, K))
    lbeta[:, T - 1] = 0.
    for t in range(T - 2, -1, -1):
        lbeta[:, t] = np.logaddexp.reduce(
            mod_tran[None] + (lbeta[:, t + 1] + ll[:, t + 1])[:, None, :], axis=2)
    return lbeta


def messages_scaled(ll, mod_init, mod_tran):
    """Scaled-domain float64 form of forward_msgs / backward_msgs / marginals (the algebra of
    SURVEY.md section 10): alpha-hat_t = normalise((alpha-hat_{t-1} P~) * b_t) with P~ =
    exp(mod_tran), b_t = exp(ll_t - max_k ll_t); beta-hat likewise; the log-domain tables follow as
    lalpha[t] = log alpha-hat_t + sum_{s<=t}(log c_s + max ll_s), lbeta[t] = log beta-hat_t +
    sum_{s>t}(log d_s + max ll_{s+1}).  One (B,K)x(K,K) product per step instead of B*K*K
    logaddexp, for the BASELINE-size parity cases where the reference-shaped recursions above
    take minutes; tests/test_oracle_golden.py holds it to them (and so to the reference-made
    fixtures) at round-off.  Returns (lalpha, lbeta)."""
    B, T, K = ll.shape
    P = np.exp(mod_tran)
    mx = np.max(ll, axis=-1)
    b = np.exp(ll - mx[..., None])
    ah = np.empty((B, T, K)); lc = np.empty((B, T))
    a = np.exp(mod_init)[None] * b[:, 0]
    for t in range(T):
        if t:
            a = ah[:, t - 1].dot(P) * b[:, t]
        c = a.sum(-1)
        ah[:, t] = a / c[:, None]
        lc[:, t] = np.log(c) + mx[:, t]
    bh = np.empty((B, T, K)); ld = np.zeros((B, T))
    bh[:, T - 1] = 1.
    for t in range(T - 2, -1, -1):
        u = (bh[:, t + 1] * b[:, t + 1]).dot(P.T)
        d = u.sum(-1)
        bh[:, t] = u / d[:, None]
        ld[:, t] = np.log(d) + mx[:, t + 1]
    with np.errstate(divide='ignore'):
        lalpha = np.log(ah) + np.cumsum(lc, axis=1)[..., None]
        lbeta = np.log(bh) + np.cumsum(ld[:, ::-1], axis=1)[:, ::-1][..., None]
    return lalpha, lbeta


def marginals(lalpha, lbeta):
    """hmmsgd_metaobs.py:516-519 / hmmbase.py:226-229."""
    v = lalpha + lbeta
    v = v - np.max(v, axis=-1, keepdims=True)
    v = np.exp(v)
    return v / np.sum(v, axis=-1, keepdims=True)


def local_update(xw, var_init, var_tran, emit, scaled=False):
    """hmmsgd_metaobs.py:487-519 on a batch of windows xw (B,T,D).
    Returns dict(ll, lalpha, lbeta, var_x).  scaled=True: the same tables through
    messages_scaled (large cases)."""
    mod_init, mod_tran = mod_params(var_init, var_tran)
    ll = lliks_gaussian(xw, emit)
    if scaled:
        lalpha, lbeta = messages_scaled(ll, mod_init, mod_tran)
    else:
        lalpha = forward_msgs(ll, mod_init, mod_tran)
        lbeta = backward_msgs(ll, mod_tran)
    return dict(ll=ll, lalpha=lalpha, lbeta=lbeta, var_x=marginals(lalpha, lbeta),
                mod_init=mod_init, mod_tran=mod_tran)


def local_lower_bound(lalpha):
    """hmmsgd_metaobs.py:257-271 (quirk Q4: sums logsumexp over *all* t)."""
    return np.sum(np.logaddexp.reduce(lalpha, axis=-1), axis=-1)


def log_Z(lalpha):
    """True log normaliser of one window = logsumexp_k lalpha[T-1,k]."""
    return np.logaddexp.reduce(lalpha[..., -1, :], axis=-1)


def exact_xi_stat(res):
    """NOT reference behaviour (the reference uses the product of marginals,
    quirk Q1).  Sum over t=1..T-1 of the true pairwise posterior
    xi_t[i,j] ~ exp(lalpha[t-1,i] + ltran[i,j] + ll[t,j] + lbeta[t,j])."""
    la, lb, ll, lt = res['lalpha'], res['lbeta'], res['ll'], res['mod_tran']
    B, T, K = ll.shape
    out = np.zeros((B, K, K))
    for t in range(1, T):
        lx = la[:, t - 1][:, :, None] + lt[None] + (ll[:, t] + lb[:, t])[:, None, :]
        lx -= lx.max(axis=(1, 2), keepdims=True)
        x = np.exp(lx)
        out += x / x.sum(axis=(1, 2), keepdims=True)
    return out


# --------------------------------------------------------------------------
# sufficient statistics
# --------------------------------------------------------------------------
def tran_stat(var_x, wrap):
    """hmmsgd_metaobs.py:876-878 (wrap=True: the index t-loff-1 = -1 at t=loff
    wraps around, quirk Q2) or hmmbatchcd.py:182-184 (wrap=False).
    var_x: (B,T,K) -> (B,K,K) product-of-marginals statistic (quirk Q1)."""
    prev = np.roll(var_x, 1, axis=1) if wrap else var_x[:, :-1]
    cur = var_x if wrap else var_x[:, 1:]
    return np.einsum('bti,btj->bij', prev, cur)


def niw_suffstats(xw, w):
    """util.py:73-83 for one state.  xw: (n,D), w: (n,) -> [xbar, neff, S, neff]."""
    tmp = w[:, None] * xw
    return [np.sum(tmp, axis=0), w.sum(), xw.T.dot(tmp), w.sum()]


def intermediate_pars(var_x, xw, maskw, prior_tran, wrap=True):
    """hmmsgd_metaobs.py:857-904 for ONE window.  var_x (T,K), xw (T,D),
    maskw (T,) bool.  Returns A_inter (K,K) (prior added per window, quirk Q5)
    and per-state [sum w x, sum w, sum w x x^T, sum w]."""
    A_inter = prior_tran + tran_stat(var_x[None], wrap)[0] - 1.
    inds = np.logical_not(maskw)
    emit_inter = [niw_suffstats(xw[inds], var_x[inds, k]) for k in range(var_x.shape[1])]
    return A_inter, emit_inter


def diag_suffstats(xw, w):
    """EXTENSION: per-dimension version of util.py:73-83 (D independent 1-D NIWs)."""
    tmp = w[:, None] * xw
    n = w.sum()
    return [np.sum(tmp, axis=0), n, np.sum(tmp * xw, axis=0), n]


# --------------------------------------------------------------------------
# NIW natural <-> moment parameters and the global steps
# --------------------------------------------------------------------------
def niw_natural(mu, sigma, kappa, nu):
    """util.py:28-37 (follows the code: eta3 = sigma + kappa mu mu^T)."""
    p = len(mu)
    return [kappa * mu, kappa, sigma + np.outer(mu, mu) * kappa, nu + 2 + p]


def niw_moment(e1, e2, e3, e4):
    """util.py:40-60 -> dict(mu, sigma, kappa, nu)."""
    p = len(e1)
    mu = e1 / e2
    kappa = e2
    return dict(mu=mu, sigma=e3 - np.outer(mu, mu) * kappa, kappa=kappa, nu=e4 - 2 - p)


def diag_natural(mu, sig, kappa, nu):
    """EXTENSION: util.py:28-37 applied per dimension with p = 1."""
    return [kappa * mu, kappa, sig + mu * mu * kappa, nu + 3.]


def diag_moment(e1, e2, e3, e4):
    mu = e1 / e2
    return dict(mu=mu, sigma=e3 - mu * mu * e2, kappa=e2, nu=e4 - 3.)


def svi_global_update(var_tran, emit, prior_emit, A_inter, emit_inter, lrate, T_full, L, S, ada_G=None):
    """hmmsgd_metaobs.py:1010-1069.  ada_G: None = plain step; a (K,K) array = the AdaGrad-like
    branch :1036-1040 (updated IN PLACE like self.ada_G; the emissions always use lrate).
    emit / prior_emit: lists of dict(mu, sigma, kappa, nu); emit_inter[k] =
    [e1,e2,e3,e4] summed over the minibatch.  Returns (var_tran_new, emit_new)."""
    bfact = (T_full - 2 * L - 1) / (2. * L * S)
    nats_old = var_tran - 1.
    if ada_G is not None:
        ada_G += nats_old ** 2
        adaMatrix = ada_G ** .25
        nats_new = (1. - 1.0 / adaMatrix) * nats_old + bfact * A_inter / adaMatrix
    else:
        nats_new = (1. - lrate) * nats_old + lrate * bfact * A_inter
    var_tran_new = nats_new + 1.
    bfact = (T_full - 2 * L - 1) / ((2. * L + 1.) * S)
    emit_new = []
    for k in range(len(emit)):
        full = np.ndim(emit[k]['sigma']) == 2
        nat, mom = (niw_natural, niw_moment) if full else (diag_natural, diag_moment)
        old = nat(emit[k]['mu'], emit[k]['sigma'], emit[k]['kappa'], emit[k]['nu'])
        pri = nat(prior_emit[k]['mu'], prior_emit[k]['sigma'], prior_emit[k]['kappa'],
                  prior_emit[k]['nu'])
        new = [(1. - lrate) * o + lrate * (p + bfact * e)
               for o, p, e in zip(old, pri, emit_inter[k])]
        emit_new.append(mom(*new))
    return var_tran_new, emit_new


def svi_minibatch_step(obs, mask, starts, T, var_tran, emit, prior_tran, prior_emit,
                       lrate, L, S=None, wrap=True, mask_ll=False, ada_G=None, scaled=False):
    """One global step of hmmsgd_metaobs.VBHMM.infer (:396-439) given the window
    start indices `starts` (window b = obs[starts[b] : starts[b]+T]).
    Returns dict with var_x (B,T,K), A_inter, emit_inter, lb, var_tran_new, emit_new."""
    T_full = obs.shape[0]
    K = var_tran.shape[0]
    S = len(starts) if S is None else S
    idx = np.asarray(starts)[:, None] + np.arange(T)[None]
    xw = obs[idx]
    mw = mask[idx] if mask is not None else np.zeros(idx.shape, bool)
    if mask_ll:                      # hmmsgd_metaobs.py:1167-1168 style NaN-masking
        xw = xw.copy()
        xw[mw] = np.nan
    var_init = stationary_init(var_tran)
    if 'alpha' in emit[0]:
        return _svi_minibatch_step_cat(xw[..., 0], mw, var_init, var_tran, emit, prior_tran, prior_emit,
                                       lrate, L, S, T_full, wrap)
    res = local_update(xw, var_init, var_tran, emit, scaled=scaled)
    A_inter = np.zeros_like(var_tran)
    emit_inter = None
    for b in range(len(starts)):
        xb = np.nan_to_num(xw[b]) if mask_ll else xw[b]
        A_i, e_i = intermediate_pars(res['var_x'][b], xb, mw[b], prior_tran, wrap) \
            if np.ndim(emit[0]['sigma']) == 2 else _intermediate_pars_diag(
                res['var_x'][b], xb, mw[b], prior_tran, wrap)
        A_inter += A_i                                           # :430
        if emit_inter is None:
            emit_inter = [[np.array(v, dtype=float) for v in e] for e in e_i]
        else:
            for k in range(K):                                   # :432-433
                for j in range(4):
                    emit_inter[k][j] = emit_inter[k][j] + e_i[k][j]
    lb = float(np.sum(local_lower_bound(res['lalpha'])))         # :436
    var_tran_new, emit_new = svi_global_update(var_tran, emit, prior_emit, A_inter,
                                               emit_inter, lrate, T_full, L, S, ada_G=ada_G)
    res.update(A_inter=A_inter, emit_inter=emit_inter, lb=lb, logZ=log_Z(res['lalpha']),
               var_tran_new=var_tran_new, emit_new=emit_new, var_init=var_init)
    return res


def _svi_minibatch_step_cat(xw, mw, var_init, var_tran, emit, prior_tran, prior_emit, lrate, L, S,
                            T_full, wrap):
    """Categorical-emission variant of svi_minibatch_step (xw: (B,T) symbols, NaN = missing)."""
    mod_init, mod_tran = mod_params(var_init, var_tran)
    ll = lliks_categorical(xw, emit)
    lalpha = forward_msgs(ll, mod_init, mod_tran)
    lbeta = backward_msgs(ll, mod_tran)
    q = marginals(lalpha, lbeta)
    B, T, K = q.shape
    C = len(emit[0]['alpha'])
    A_inter = np.zeros_like(var_tran)
    counts = np.zeros((K, C))
    for b in range(B):
        A_inter += prior_tran + tran_stat(q[b][None], wrap)[0] - 1.
        inds = np.logical_not(mw[b]) & ~np.isnan(xw[b])
        for k in range(K):
            counts[k] += cat_suffstats(xw[b][inds], q[b][inds, k], C)
    bA = (T_full - 2 * L - 1) / (2. * L * S)
    bE = (T_full - 2 * L - 1) / ((2. * L + 1.) * S)
    var_tran_new = (1. - lrate) * (var_tran - 1.) + lrate * bA * A_inter + 1.
    emit_new = [dict(alpha=cat_global_update(emit[k]['alpha'], prior_emit[k]['alpha'], counts[k], B,
                                             lrate, bE)) for k in range(K)]
    return dict(ll=ll, lalpha=lalpha, lbeta=lbeta, var_x=q, mod_init=mod_init, mod_tran=mod_tran,
                A_inter=A_inter, counts=counts, lb=float(np.sum(local_lower_bound(lalpha))),
                logZ=log_Z(lalpha), var_tran_new=var_tran_new, emit_new=emit_new, var_init=var_init)


def _intermediate_pars_diag(var_x, xw, maskw, prior_tran, wrap):
    A_inter = prior_tran + tran_stat(var_x[None], wrap)[0] - 1.
    inds = np.logical_not(maskw)
    return A_inter, [diag_suffstats(xw[inds], var_x[inds, k]) for k in range(var_x.shape[1])]


# --------------------------------------------------------------------------
# batch coordinate ascent (BASELINE config 1)
# --------------------------------------------------------------------------
def niw_posterior(prior, xw, w):
    """pybasicbayes/distributions.py:240-276 (+:324-329): conjugate NIW update from
    weighted (n, xbar, centred scatter); keeps the prior when n <= weps."""
    n = w.sum()
    if not n > WEPS:
        return dict(mu=prior['mu'], sigma=prior['sigma'], kappa=prior['kappa'], nu=prior['nu'])
    xbar = np.dot(w, xw) / n
    c = xw - xbar
    sumsq = np.dot(c.T, w[:, None] * c)
    k0, m0 = prior['kappa'], prior['mu']
    return dict(mu=k0 / (k0 + n) * m0 + n / (k0 + n) * xbar,
                sigma=prior['sigma'] + sumsq + k0 * n / (k0 + n) * np.outer(xbar - m0, xbar - m0),
                kappa=k0 + n, nu=prior['nu'] + n)


def batch_cavi_step(obs, mask, var_init, var_tran, emit, prior_init, prior_tran, prior_emit):
    """One iteration of hmmbatchcd.VBHMM.infer (:135-141): base local_update
    (hmmbase.py:201-229) then global_update (hmmbatchcd.py:172-189)."""
    res = local_update(obs[None], var_init, var_tran, emit)
    q = res['var_x'][0]
    new_init = prior_init + q[0]
    new_tran = prior_tran + tran_stat(q[None], wrap=False)[0]
    inds = np.logical_not(mask) if mask is not None else np.ones(len(obs), bool)
    new_emit = [niw_posterior(prior_emit[k], obs[inds], q[inds, k]) for k in range(q.shape[1])]
    res.update(var_init_new=new_init, var_tran_new=new_tran, emit_new=new_emit,
               lZ=float(local_lower_bound(res['lalpha'])[0]))
    return res


# --------------------------------------------------------------------------
# batch natural gradient (hmmbatchsgd.py)
# --------------------------------------------------------------------------
def batch_sgd_step(obs, mask, var_init, var_tran, emit, prior_init, prior_tran, prior_emit, lrate):
    """One iteration of hmmbatchsgd.VBHMM.infer (:160-168): base local_update (hmmbase.py:201-229) on
    the NaN-masked observations (:148-149) then global_update (:202-259): the natural parameters move
    a step `lrate` towards those of the conjugate full-batch update."""
    xo = obs.copy()
    if mask is not None:
        xo[mask] = np.nan
    res = local_update(xo[None], var_init, var_tran, emit)
    q = res['var_x'][0]
    new_init = prior_init + q[0]                                        # :216
    tran_mf = prior_tran + tran_stat(q[None], wrap=False)[0]            # :222-225
    new_tran = (1. - lrate) * (var_tran - 1.) + lrate * (tran_mf - 1.) + 1.
    inds = np.logical_not(mask) if mask is not None else np.ones(len(obs), bool)
    new_emit = []
    for k in range(q.shape[1]):
        post = niw_posterior(prior_emit[k], obs[inds], q[inds, k])      # util.NIW_meanfield
        nt = niw_natural(post['mu'], post['sigma'], post['kappa'], post['nu'])
        no = niw_natural(emit[k]['mu'], emit[k]['sigma'], emit[k]['kappa'], emit[k]['nu'])
        new_emit.append(niw_moment(*[(1. - lrate) * o + lrate * t for o, t in zip(no, nt)]))
    res.update(var_init_new=new_init, var_tran_new=new_tran, emit_new=new_emit)
    return res


# --------------------------------------------------------------------------
# adaptive window machinery
# --------------------------------------------------------------------------
def get_local_messages(obs, ind, halflength, var_init, var_tran, emit):
    """hmmsgd_metaobs.py:663-700: marginals of the window [ind-halflength, ind+halflength] with
    the GIVEN var_init (whatever self.var_init holds, not recomputed)."""
    xw = obs[ind - halflength: ind + halflength + 1][None]
    return local_update(xw, var_init, var_tran, emit)['var_x'][0]


def select_L(obs, indices, var_init, var_tran, emit, epsilon=1e-5, minHalfL=1, Lincrement=1, Lcutoff=1000):
    """hmmsgd_metaobs.py:521-545 (non-averaged branch) for given centre indices."""
    T = obs.shape[0]
    maxL = -1
    for ind in indices:
        q_diff = np.finfo(np.float64).max
        L = minHalfL
        q_old = get_local_messages(obs, ind, minHalfL, var_init, var_tran, emit)[minHalfL]
        while True:
            if ind - L < 1 + Lincrement or ind + L + Lincrement + 1 > T or L > Lcutoff:
                break
            if q_diff < epsilon:
                break
            L += Lincrement
            q_new = get_local_messages(obs, ind, L, var_init, var_tran, emit)[L]
            q_diff = np.sum(np.abs(q_new - q_old))
            q_old = q_new
        maxL = max(maxL, L)
    return maxL


def select_buffer(obs, indices, var_init, var_tran, emit, epsilon=1e-5, halfL=10, Lincrement=1, Lcutoff=1000):
    """hmmsgd_metaobs.py:579-625 (non-averaged branch) for given centre indices."""
    T = obs.shape[0]
    maxL = -1
    for ind in indices:
        dl = dr = np.finfo(np.float64).max
        bufferL = halfL
        v = get_local_messages(obs, ind, halfL, var_init, var_tran, emit)
        ql, qr = v[bufferL - halfL], v[bufferL + halfL]
        while True:
            if ind - bufferL < 1 + Lincrement or ind + bufferL + Lincrement + 1 > T or bufferL > Lcutoff:
                break
            if dl < epsilon and dr < epsilon:
                break
            bufferL += Lincrement
            v = get_local_messages(obs, ind, bufferL, var_init, var_tran, emit)
            nl, nr = v[bufferL - halfL], v[bufferL + halfL]
            dl, dr = np.sum(np.abs(nl - ql)), np.sum(np.abs(nr - qr))
            ql, qr = nl, nr
        maxL = max(maxL, bufferL)
    return maxL


def buffered_stats(obs, mask, starts, bufferL, L, var_tran, emit, prior_tran):
    """local_update on the buffered windows [start, start+2*bufferL] then intermediate_pars_buffer
    (hmmsgd_metaobs.py:932-1008): statistics from var_x[bufferL-L : bufferL+L+1] with the wrap-around
    inside that slice.  Returns dict(var_x (B,Tbuf,K), A_inter, emit_inter, lb)."""
    Tb = 2 * bufferL + 1
    idx = np.asarray(starts)[:, None] + np.arange(Tb)[None]
    xw = obs[idx]
    mw = mask[idx] if mask is not None else np.zeros(idx.shape, bool)
    res = local_update(xw, stationary_init(var_tran), var_tran, emit)
    A_inter = np.zeros_like(var_tran)
    emit_inter = None
    lo, hi = bufferL - L, bufferL + L + 1
    for b in range(len(starts)):
        A_i, e_i = intermediate_pars(res['var_x'][b][lo:hi], xw[b][lo:hi], mw[b][lo:hi], prior_tran, True)
        A_inter += A_i
        emit_inter = e_i if emit_inter is None else [[u + v for u, v in zip(a, c)] for a, c in zip(emit_inter, e_i)]
    return dict(var_x=res['var_x'], A_inter=A_inter, emit_inter=emit_inter,
                lb=float(np.sum(local_lower_bound(res['lalpha']))))


# --------------------------------------------------------------------------
# predictive log-probability of the held-out rows (hmmsgd_metaobs.py:1086-1205)
# --------------------------------------------------------------------------
def full_local_update(obs, mask, var_init, var_tran, emit):
    """hmmsgd_metaobs.py:1147-1205: marginals of the whole series with the masked rows NaN-ed
    (they carry no evidence: ll = 0 through nan_to_num)."""
    xo = obs.copy()
    xo[np.asarray(mask, bool)] = np.nan
    return local_update(xo[None], var_init, var_tran, emit)['var_x'][0]


def pred_logprob(var_x, obs, mask, emit):
    """hmmsgd_metaobs.py:1111-1119 / :1139-1145: mean over the masked rows of
    logsumexp_k(log(var_x + eps) + ELL_k(x)) with the TRUE observations; None without masked rows."""
    m = np.asarray(mask, bool)
    if not m.any():
        return None
    ll = lliks_gaussian(obs[m][None], emit)[0]
    return float(np.mean(np.logaddexp.reduce(np.log(var_x[m] + EPS) + ll, axis=1)))


# --------------------------------------------------------------------------
# forward-filter backward-sampling (hmm_fast.pyx:43-124)
# --------------------------------------------------------------------------
def ffbs_tables(obs, var_init, var_tran, emit):
    """The distribution hmm_fast.FFBS samples from: forward table with log(A + eps) transition weights
    (hmm_fast.pyx:82-100; DBL_EPSILON) and the exact marginals / pairwise marginals of the sampled
    path, P(z_t), P(z_t, z_{t+1}), obtained by a backward pass with the same weights."""
    eps = np.finfo(np.float64).eps
    mod_init = digamma(var_init + eps) - digamma(np.sum(var_init) + eps)
    ltran = np.log(var_tran + eps)
    ll = lliks_gaussian(obs[None], emit)
    lalpha = forward_msgs(ll, mod_init, ltran)[0]
    lbeta = backward_msgs(ll, ltran)[0]
    marg = marginals(lalpha[None], lbeta[None])[0]
    T, K = marg.shape
    pair = np.empty((T - 1, K, K))
    for t in range(T - 1):
        lx = lalpha[t][:, None] + ltran + (ll[0, t + 1] + lbeta[t + 1])[None, :]
        x = np.exp(lx - lx.max())
        pair[t] = x / x.sum()
    return lalpha, marg, pair
