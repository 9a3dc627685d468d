"""Debug probe of the tensor-core emission kernel (emit_tc.cuh): marginals and statistics of the same
minibatch with k_emit_tc (default) and with the float64 kernel (SVIHMM_KEEP_LOCALS keeps it), error
statistics against the float64 oracle for small cases, phase timings.  No assertions."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pysvihmm_b200 import _lib as L  # noqa: E402
from pysvihmm_b200.engine import EStepEngine  # noqa: E402
from tests.helpers import make_random_problem, pack_emit_np  # noqa: E402

cases = [tuple(int(v) for v in c.split(",")) for c in sys.argv[1:]] or [(64, 32, 300, 5), (64, 32, 1024, 64)]
for K, D, T, B in cases:
    print("case", K, D, T, B, flush=True)
    p = make_random_problem(seed=K * 1000 + T, K=K, D=D, T_full=max(8 * T, 4000), kind="niw_full", miss=0.05, sep=0.4)
    obs = p["obs"].astype(np.float32).astype(np.float64)
    obs[17, 3] = np.nan
    starts = np.random.RandomState(5).randint(0, obs.shape[0] - T + 1, B)
    starts[0] = 0
    eng = EStepEngine(K, D, "niw_full")
    eng.set_series(obs, p["mask"], dtype="f32")
    eng.set_prior(p["prior_tran"], pack_emit_np(p["prior_emit"]))
    eng.set_globals(p["var_tran"], pack_emit_np(p["emit"]))
    eng.set_profiling(True)
    out = {}
    for name, fl in (("tc", 0), ("f64", L.KEEP_LOCALS)):
        for rep in range(3):
            vx, stats = eng.estep(starts, T, flags=L.WRAP | L.ADD_PRIOR | fl)
        out[name] = (vx.cpu().numpy().copy(), eng.unpack_stats(stats))
        print("  %-4s phases ms (3 reps): %s" % (name, {k: round(v[0], 3) for k, v in eng.phase_ms().items()}), flush=True)
    q1, s1 = out["tc"]; q0, s0 = out["f64"]
    d = np.abs(q1 - q0)
    print("  max |q_tc - q_f64| %.3e  excess over 1e-5 q + 2e-7: %.3e  nan %d  rowsum err %.2e" % (
        d.max(), (d - 1e-5 * q0 - 2e-7).max(), int(np.isnan(q1).sum()), np.abs(q1.sum(-1) - 1).max()), flush=True)
    w = np.unravel_index(np.argmax(d), d.shape)
    print("  worst at (b, t, k) =", w, "q_tc", q1[w], "q_f64", q0[w], flush=True)
    for key in ("A", "n", "sx", "sxx"):
        print("  %s rel %.3e" % (key, np.abs(s1[key] - s0[key]).max() / np.abs(s0[key]).max()), end="")
    print("  logZ %.10g vs %.10g" % (s1["logZ"], s0["logZ"]), flush=True)
    if B * T <= 40000:
        from oracle import svihmm_oracle as O
        r = O.svi_minibatch_step(obs, p["mask"], starts, T, p["var_tran"], p["emit"], p["prior_tran"],
                                 p["prior_emit"], 0.37, max(T // 2, 1), wrap=True, scaled=True)
        for name in ("tc", "f64"):
            e = np.abs(out[name][0] - r["var_x"])
            print("  %-4s vs oracle: max %.3e  excess %.3e" % (name, e.max(), (e - 1e-5 * r["var_x"] - 2e-7).max()), flush=True)
    eng.close()
