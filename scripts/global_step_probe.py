import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from pysvihmm_b200.engine import EStepEngine, pack_emit_dicts
cfg = bench.CONFIGS[os.environ.get("DBG_CONFIG", "c2")]
K, D, T, B, kind = cfg["K"], cfg["D"], cfg["T"], int(os.environ.get("DBG_B", cfg["B"])), cfg["kind"]
obs, mus = bench.synthetic_series(K, D, 1 << 18, 1)
vt, em, pr = bench.globals_for(K, D, kind, mus, 1)
eng = EStepEngine(K, D, kind)
eng.set_series(torch.from_numpy(obs).cuda())
eng.set_prior(np.ones((K, K)), pack_emit_dicts(pr))
eng.set_globals(vt, pack_emit_dicts(em))
st = torch.randint(0, (1 << 18) - T, (B,))
_, stats = eng.estep(st, T, flags=3)
for i in range(4):
    eng.global_update(stats, 0.3, 1.0, 1.0)
torch.cuda.synchronize()
