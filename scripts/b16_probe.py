"""Debug probe of the batched K <= 16 path: runs listed (K, D, T, B) cases and prints error statistics
against the oracle (no assertions)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import svihmm_oracle as O  # noqa: E402
from pysvihmm_b200 import _lib as L  # noqa: E402
from pysvihmm_b200.engine import EStepEngine  # noqa: E402
from tests.helpers import make_random_problem, pack_emit_np  # noqa: E402

cases = [tuple(int(v) for v in c.split(",")) for c in sys.argv[1:]] or [(16, 8, 257, 37)]
for K, D, T, B in cases:
    print("case", K, D, T, B, flush=True)
    p = make_random_problem(seed=K * 1000 + T, K=K, D=D, T_full=max(4 * T, 300), kind="niw_diag", miss=0.1)
    starts = np.random.RandomState(5).randint(0, p["obs"].shape[0] - T + 1, B)
    eng = EStepEngine(K, D, "niw_diag")
    eng.set_series(p["obs"], p["mask"], dtype="f64")
    eng.set_prior(p["prior_tran"], pack_emit_np(p["prior_emit"]))
    eng.set_globals(p["var_tran"], pack_emit_np(p["emit"]))
    t0 = time.time()
    vx, stats = eng.estep(starts, T, flags=L.WRAP | L.ADD_PRIOR)
    q = vx.cpu().numpy()
    print("  estep + copy %.3fs" % (time.time() - t0), flush=True)
    r = O.svi_minibatch_step(p["obs"], p["mask"], starts, T, p["var_tran"], p["emit"], p["prior_tran"],
                             p["prior_emit"], 0.5, max(T // 2, 1), wrap=True, scaled=True)
    err = np.abs(q - r["var_x"])
    print("  max |q - ref| %.3e, max excess over 1e-5 q + 2e-7: %.3e, nan %d" % (
        err.max(), (err - 1e-5 * r["var_x"] - 2e-7).max(), int(np.isnan(q).sum())), flush=True)
    s = eng.unpack_stats(stats)
    A = O.tran_stat(r["var_x"], True).sum(0) + B * (p["prior_tran"] - 1.)
    print("  A rel err %.3e  logZ rel %.3e  q4 rel %.3e  B %d" % (
        np.abs(s["A"] - A).max() / np.abs(A).max(), abs(s["logZ"] - r["logZ"].sum()) / abs(r["logZ"].sum()),
        abs(s["lb_q4"] - r["lb"]) / abs(r["lb"]), s["B"]), flush=True)
    n = np.array([e[1] for e in r["emit_inter"]]); sx = np.array([e[0] for e in r["emit_inter"]]); sxx = np.array([e[2] for e in r["emit_inter"]])
    print("  n rel %.3e  sx rel %.3e  sxx rel %.3e" % (np.abs(s["n"] - n).max() / n.max(), np.abs(s["sx"] - sx).max() / np.abs(sx).max(),
                                                    np.abs(s["sxx"] - sxx).max() / np.abs(sxx).max()), flush=True)
    eng.close()
