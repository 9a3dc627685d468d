#!/usr/bin/env python
"""Recipe that makes the reference's own CPU path runnable on Python 3.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this.

The reference (dillonalaird/pysvihmm @ fed7cff1, Python 2, numpy<1.24) cannot
be imported by this image's Python 3.12 / numpy 2.3.  This script reads the
reference sources *where they lie* under /root/reference, applies the purely
mechanical Python-2 -> Python-3 patches listed in PATCHES below (no arithmetic
is touched) and writes the result into ``oracle/_ref/`` (git-ignored: the
patched copy is a build artefact, never committed; it still travels to the GPU
box with the snapshot so ``bench.py --impl reference`` can time it there).

The one third-party dependency that is absent from the reference tree,
``pybasicbayes/util`` (= mattjj/pymattutil @ 9e59824b, an un-vendored nested
submodule, see /root/reference/.SUBMODULES.json), is replaced by a small shim
written here from the published definitions of the handful of helpers that
are imported.  None of them is on the E-step arithmetic path (SURVEY.md
section 8c): they are reached only from ``Gaussian.resample`` (random
initialisation, which every parity test avoids by passing explicit mu/sigma)
and ``Gaussian.get_vlb`` (ELBO diagnostic).

Usage:  python oracle/build_ref.py [--src /root/reference] [--dst oracle/_ref]
"""
import argparse
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

FILES = [
    "hmmbase.py",
    "hmmsgd_metaobs.py",
    "hmmbatchcd.py",
    "hmmbatchsgd.py",
    "util.py",
    "gen_synthetic.py",
    "munkres.py",
    "pybasicbayes/abstractions.py",
    "pybasicbayes/distributions.py",
]

# (description, regex, replacement) applied to every file, in order.
PATCHES = [
    ("print statement -> function",
     re.compile(r"^(\s*)print (?!\()(.*)$", re.M), r"\1print(\2)"),
    ("xrange -> range", re.compile(r"\bxrange\b"), "range"),
    ("cPickle -> pickle", re.compile(r"import cPickle as pkl"), "import pickle as pkl"),
    (".iteritems/.itervalues", re.compile(r"\.iter(items|values)\(\)"), r".\1()"),
    ("long type", re.compile(r"\(int,long,float,complex\)"), "(int,float,complex)"),
    ("np.float_ removed in numpy 2", re.compile(r"np\.float_\b"), "np.float64"),
    ("3-arg MethodType (unbound methods are gone)",
     re.compile(r"types\.MethodType\(hmm_fast\.FFBS, None, VariationalHMMBase\)"),
     "hmm_fast.FFBS"),
    ("inner1d moved/removed",
     re.compile(r"^from numpy\.core\.umath_tests import inner1d$", re.M),
     "inner1d = lambda a, b: np.einsum('ij,ij->i', a, b)"),
    ("scipy.weave is gone", re.compile(r"^import scipy\.weave$", re.M), "import scipy.linalg"),
    ("matplotlib is not installed (plots only)",
     re.compile(r"^import matplotlib\.pyplot as plt$", re.M),
     "plt = None"),
    ("implicit relative imports inside pybasicbayes",
     re.compile(r"^from abstractions import", re.M), "from .abstractions import"),
    ("implicit relative imports inside pybasicbayes",
     re.compile(r"^from util\.stats import", re.M), "from .util.stats import"),
    ("ragged np.array([...]) needs dtype=object on numpy>=1.24 (util.py:14)",
     re.compile(r"return np\.array\(\[np\.zeros\(p\), 0\., np\.zeros\(\(p,p\)\), 0\]\)"),
     "return np.array([np.zeros(p), 0., np.zeros((p,p)), 0], dtype=object)"),
    ("ragged np.array (util.py:25)",
     re.compile(r"return np\.array\(\[mu, sigma, kappa, nu\]\)"),
     "return np.array([mu, sigma, kappa, nu], dtype=object)"),
    ("ragged np.array (util.py:37)",
     re.compile(r"return np\.array\(\[kappa \* mu, kappa, eta3, nu \+ 2 \+ p\]\)"),
     "return np.array([kappa * mu, kappa, eta3, nu + 2 + p], dtype=object)"),
    ("ragged np.array (util.py:70)",
     re.compile(r"return np\.array\(\[mu_mf, sigma_mf, kappa_mf, nu_mf\]\)"),
     "return np.array([mu_mf, sigma_mf, kappa_mf, nu_mf], dtype=object)"),
    ("ragged np.array (util.py:83)",
     re.compile(r"return np\.array\(\[xbar, neff, S, neff\]\)"),
     "return np.array([xbar, neff, S, neff], dtype=object)"),
    ("tuple==(None,None) with arrays raises (distributions.py:211,286)",
     re.compile(r"\(mu,sigma\) == \(None,None\)"), "(mu is None and sigma is None)"),
    ("same, attribute form",
     re.compile(r"\(self\.mu,self\.sigma\) == \(None,None\)"),
     "(self.mu is None and self.sigma is None)"),
    ("`None not in (arrays)` compares arrays elementwise (distributions.py:211)",
     re.compile(r"None not in \(mu_0,sigma_0,kappa_0,nu_0\)"),
     "all(v is not None for v in (mu_0,sigma_0,kappa_0,nu_0))"),
    ("negative float slice index (util.py:204)",
     re.compile(r"mask\[-nmiss:\]"), "mask[-int(nmiss):]"),
]

HMM_FAST_STUB = '''"""Stub for the Cython FFBS sampler (hmm_fast.pyx): not on the E-step path."""
def FFBS(self, var_init, lalpha_init=None):
    raise NotImplementedError("hmm_fast.FFBS is a Cython module; not built in oracle/_ref")
'''

UTIL_STATS_SHIM = '''"""Shim for the absent submodule pybasicbayes/util (mattjj/pymattutil @ 9e59824b).

Written from the published definitions; cannot be verified against the source
here (it is not in /root/reference).  Off the E-step arithmetic path.
"""
import numpy as np
import scipy.special as special
import scipy.linalg


def getdatasize(data):
    if isinstance(data, np.ndarray):
        return data.shape[0]
    if isinstance(data, list):
        return sum(getdatasize(d) for d in data)
    return 1


def getdatadimension(data):
    if isinstance(data, np.ndarray):
        return data.shape[1] if data.ndim > 1 else 1
    return getdatadimension(data[0])


def gi(data):
    out = (np.isnan(np.atleast_2d(data)).sum(1) == 0).ravel()
    return out if len(out) != 1 else out[0]


def atleast_2d(data):
    return data if data.ndim > 1 else data.reshape((-1, 1))


def flattendata(data):
    if isinstance(data, np.ndarray):
        return data
    return np.concatenate(data)


def combinedata(datas):
    out = []
    for d in datas:
        if isinstance(d, np.ndarray):
            out.append(d)
        elif isinstance(d, list):
            out.extend(d)
        else:
            out.append(np.atleast_1d(d))
    return out


def sample_discrete(distn, size=[], dtype=np.int32):
    distn = np.atleast_1d(distn)
    cumvals = np.cumsum(distn)
    return np.sum(np.array(np.random.random(size))[..., None] * cumvals[-1] > cumvals,
                  axis=-1, dtype=dtype)


def sample_discrete_from_log(p_log, axis=0, dtype=np.int32):
    cumvals = np.exp(p_log - np.expand_dims(p_log.max(axis), axis)).cumsum(axis)
    thesize = np.array(p_log.shape)
    thesize[axis] = 1
    randvals = np.random.random(size=thesize) * \\
        np.reshape(cumvals[tuple([slice(None) if i is not axis else -1
                                  for i in range(p_log.ndim)])], thesize)
    return np.sum(randvals > cumvals, axis=axis, dtype=dtype)


def sample_pareto(x_m, alpha):
    return x_m + np.random.pareto(alpha)


def sample_invwishart(lmbda, dof):
    n = lmbda.shape[0]
    chol = np.linalg.cholesky(lmbda)
    if (dof <= 81 + n) and (dof == np.round(dof)):
        x = np.random.randn(int(dof), n)
    else:
        x = np.diag(np.sqrt(np.atleast_1d(np.random.chisquare(dof - np.arange(n)))))
        x[np.triu_indices_from(x, 1)] = np.random.randn(n * (n - 1) // 2)
    R = np.linalg.qr(x, 'r')
    T = scipy.linalg.solve_triangular(R.T, chol.T, lower=True).T
    return np.dot(T, T.T)


def sample_niw(mu, lmbda, kappa, nu):
    lmbda = sample_invwishart(lmbda, nu)
    mu = np.random.multivariate_normal(mu, lmbda / kappa)
    return mu, lmbda


def invwishart_entropy(sigma, nu, chol=None):
    D = sigma.shape[0]
    chol = np.linalg.cholesky(sigma) if chol is None else chol
    Elogdetlmbda = special.digamma((nu - np.arange(D)) / 2).sum() + D * np.log(2) \\
        - 2 * np.log(chol.diagonal()).sum()
    return invwishart_log_partitionfunction(sigma, nu, chol) - (nu - D - 1) / 2 * Elogdetlmbda \\
        + nu * D / 2


def invwishart_log_partitionfunction(sigma, nu, chol=None):
    D = sigma.shape[0]
    chol = np.linalg.cholesky(sigma) if chol is None else chol
    return -1 * (nu * np.log(chol.diagonal()).sum()
                 - (nu * D / 2 * np.log(2) + D * (D - 1) / 4 * np.log(np.pi)
                    + special.gammaln((nu - np.arange(D)) / 2).sum()))


def multivariate_t_loglik(y, nu, mu, lmbda):
    d = len(mu)
    yc = np.array(y - mu, ndmin=2)
    L = np.linalg.cholesky(lmbda)
    ys = scipy.linalg.solve_triangular(L, yc.T, overwrite_b=True, lower=True)
    return special.gammaln((nu + d) / 2.) - special.gammaln(nu / 2.) \\
        - (d / 2.) * np.log(nu * np.pi) - np.log(L.diagonal()).sum() \\
        - (nu + d) / 2. * np.log1p(1. / nu * np.einsum('ij,ij->j', ys, ys))
'''


# hmm_fast.pyx (the reference's only native binding): mechanical patches for Cython 3 / numpy 2
PYX_PATCHES = [
    ("xrange -> range", re.compile(r"\bxrange\b"), "range"),
    ("np.int_t is gone from numpy.pxd", re.compile(r"np\.int_t\b"), "np.int64_t"),
    ("np.int_ -> fixed width to match the buffer type", re.compile(r"dtype=np\.int_\b"), "dtype=np.int64"),
    ("array == None", re.compile(r"lalpha_init == None"), "lalpha_init is None"),
]


def build_hmm_fast(src, dst, verbose=True):
    """Cythonize and compile the reference's FFBS sampler (hmm_fast.pyx) into dst; returns the path
    of the extension module or None (the stub hmm_fast.py then stays in charge)."""
    import subprocess
    import sysconfig
    try:
        import numpy
        text = open(os.path.join(src, "hmm_fast.pyx")).read()
        for _, rx, rep in PYX_PATCHES:
            text = rx.sub(rep, text)
        pyx = os.path.join(dst, "hmm_fast.pyx")
        with open(pyx, "w") as f:
            f.write("# GENERATED by oracle/build_ref.py from hmm_fast.pyx -- do not commit\n" + text)
        subprocess.check_call([sys.executable, "-m", "cython", "-3", pyx, "-o", os.path.join(dst, "hmm_fast.c")],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        so = os.path.join(dst, "hmm_fast" + sysconfig.get_config_var("EXT_SUFFIX"))
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-w", "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
                               "-I", sysconfig.get_paths()["include"], "-I", numpy.get_include(),
                               os.path.join(dst, "hmm_fast.c"), "-o", so])
        if verbose:
            print("hmm_fast.pyx                       cythonized and compiled -> %s" % os.path.basename(so))
        return so
    except Exception as e:                      # no compiler / cython: the E-step path does not need it
        if verbose:
            print("hmm_fast.pyx                       NOT built (%s); stub kept" % e)
        return None


def patch_text(text):
    applied = []
    for desc, rx, rep in PATCHES:
        new, n = rx.subn(rep, text)
        if n:
            applied.append((desc, n))
            text = new
    return text, applied


def build(src, dst, verbose=True):
    if not os.path.isdir(src):
        raise SystemExit("reference tree %s not found (only available in the build container)" % src)
    os.makedirs(os.path.join(dst, "pybasicbayes", "util"), exist_ok=True)
    for rel in FILES:
        with open(os.path.join(src, rel), "r") as f:
            text = f.read()
        text, applied = patch_text(text)
        out = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        with open(out, "w") as f:
            f.write("# GENERATED by oracle/build_ref.py from %s -- do not commit\n" % rel)
            f.write(text)
        if verbose:
            print("%-34s %s" % (rel, ", ".join("%s x%d" % a for a in applied) or "(verbatim)"))
    with open(os.path.join(dst, "hmm_fast.py"), "w") as f:
        f.write(HMM_FAST_STUB)
    # package files: the reference's pybasicbayes/__init__.py imports models.py
    # (matplotlib, mixture models) which is not on the path -> minimal __init__.
    with open(os.path.join(dst, "pybasicbayes", "__init__.py"), "w") as f:
        f.write("# generated: only abstractions/distributions are needed on the E-step path\n")
    with open(os.path.join(dst, "pybasicbayes", "util", "__init__.py"), "w") as f:
        f.write("")
    with open(os.path.join(dst, "pybasicbayes", "util", "stats.py"), "w") as f:
        f.write(UTIL_STATS_SHIM)
    with open(os.path.join(dst, "__init__.py"), "w") as f:
        f.write("")
    build_hmm_fast(src, dst, verbose)
    return dst


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ap.add_argument("--dst", default=os.path.join(HERE, "_ref"))
    a = ap.parse_args()
    build(a.src, a.dst)
    sys.path.insert(0, a.dst)
    import hmmsgd_metaobs, hmmbatchcd  # noqa: F401  (import check)
    print("oracle/_ref built and importable")
