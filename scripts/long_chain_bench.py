"""Time one E-step over ONE long chain (B = 1, T = 1e6, K = 16): the block-parallel scan (default) or,
with SVIHMM_NO_SCAN=1, the sequential kernels.  Prints device milliseconds per phase."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pysvihmm_b200 import _lib as L  # noqa: E402
from pysvihmm_b200.engine import EStepEngine, pack_emit_dicts  # noqa: E402
import bench  # noqa: E402

K, D, T = 16, 8, int(os.environ.get("LC_T", 1000000))
obs, mus = bench.synthetic_series(K, D, T, seed=3, sep=1.0)
var_tran, emit, prior = bench.globals_for(K, D, "niw_diag", mus, seed=3)
eng = EStepEngine(K, D, "niw_diag")
eng.set_series(torch.from_numpy(obs).cuda())
eng.set_prior(np.ones((K, K)), pack_emit_dicts(prior))
eng.set_globals(var_tran, pack_emit_dicts(emit))
vx = torch.empty((1, T, K), dtype=torch.float32, device="cuda")
st = eng.new_stats()
for _ in range(2):
    eng.estep([0], T, flags=L.MASK_LL, var_x=vx, stats=st)
torch.cuda.synchronize()
eng.set_profiling(True); eng.phase_ms()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 3
for _ in range(n):
    eng.estep([0], T, flags=L.MASK_LL, var_x=vx, stats=st)
e1.record(); torch.cuda.synchronize()
ph = eng.phase_ms()
print("scan" if not os.environ.get("SVIHMM_NO_SCAN") else "sequential", "T=%d K=%d: %.3f ms per E-step;" % (T, K, e0.elapsed_time(e1) / n),
      {k: round(v[0] / n, 3) for k, v in ph.items()}, "logZ", float(st[-4].item()))
