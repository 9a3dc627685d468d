"""Debug: tables of the block-parallel scan against the sequential kernels (same process)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pysvihmm_b200 import _lib as L
from pysvihmm_b200.engine import EStepEngine
from tests.helpers import make_random_problem, pack_emit_np
K, D, T, B, kind = 16, 4, 5000, 4, "niw_full"
p = make_random_problem(seed=K * 7 + D, K=K, D=D, T_full=B * T + 100, kind=kind, miss=0.05, sep=0.5)
starts = np.arange(B) * T + 50
out = {}
for name, minT in (("scan", 4096), ("seq", 0)):
    eng = EStepEngine(K, D, kind)
    eng.set_tuning(L.TUNE_SCAN_MIN_T, minT)
    eng.set_series(p["obs"], p["mask"], dtype="f64")
    eng.set_prior(p["prior_tran"], pack_emit_np(p["prior_emit"]))
    eng.set_globals(p["var_tran"], pack_emit_np(p["emit"]))
    vx, stats = eng.estep(starts, T, flags=L.MASK_LL | L.ADD_PRIOR, keep_locals=True)
    loc = eng.get_locals(B, T)
    loc["q"] = vx.cpu().numpy()
    out[name] = loc
    eng.close()
for key in ("alpha", "cs", "beta", "sb", "q"):
    a, b = out["scan"][key].astype(np.float64), out["seq"][key].astype(np.float64)
    rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
    idx = np.unravel_index(np.argmax(np.abs(a - b)), a.shape)
    print(key, "max abs diff %.3e at %s (scan %.6e seq %.6e), median rel %.2e" % (np.abs(a - b).max(), idx, a[idx], b[idx], np.median(rel)))
sb_s, sb_q = out["scan"]["sb"][0], out["seq"]["sb"][0]
bad = np.nonzero(np.abs(sb_s - sb_q) > 1e-4 * np.abs(sb_q))[0]
print("rows with sb off by > 1e-4 rel (window 0):", bad[:40], "count", len(bad))
if len(bad):
    print("ratios", (sb_s[bad] / sb_q[bad])[:20])
