"""torchrun --nproc-per-node N scripts/peer_allreduce_check.py: the fused peer all-reduce + update
(svihmm_global_update_peers) against NCCL all-reduce + svihmm_global_update, same inputs."""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch, torch.distributed as dist
from pysvihmm_b200 import _lib as L
from pysvihmm_b200.engine import EStepEngine, pack_emit_dicts
from pysvihmm_b200.sharding import PeerExchange, allreduce_stats
from tests.helpers import make_random_problem

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
for K, D, kind, T in [(16, 8, "niw_diag", 64), (5, 3, "niw_full", 40), (40, 4, "niw_full", 33)]:
    p = make_random_problem(seed=K, K=K, D=D, T_full=2000, kind=kind, miss=0.05)
    rs = np.random.RandomState(7)
    engs = []
    for _ in range(2):
        e = EStepEngine(K, D, kind, device=local)
        e.set_series(p["obs"], p["mask"], dtype="f64")
        e.set_prior(p["prior_tran"], pack_emit_dicts(p["prior_emit"]))
        e.set_globals(p["var_tran"], pack_emit_dicts(p["emit"]))
        engs.append(e)
    ref, new = engs
    px = PeerExchange(new, dist)
    st = ref.new_stats()
    for it in range(5):
        starts = rs.randint(0, 2000 - T, (world, 9))[rank]
        ref.estep(starts, T, flags=L.WRAP | L.ADD_PRIOR, stats=st, want_var_x=False)
        allreduce_stats(st, dist)
        ref.global_update(st, (it + 1.) ** -0.7, 2.0, 1.5)
        st2 = new.new_stats()
        new.estep(starts, T, flags=L.WRAP | L.ADD_PRIOR, stats=st2, want_var_x=False)
        px.global_update(st2, (it + 1.) ** -0.7, 2.0, 1.5)
        a, b = st.cpu().numpy(), px.reduced_stats().cpu().numpy()
        ok &= bool(np.allclose(a, b, rtol=1e-12, atol=1e-12))
    for x, y in zip(ref.get_globals(), new.get_globals()):
        ok &= bool(np.allclose(x, y, rtol=1e-10, atol=1e-12))
    # replicas in lock-step: identical bits on every rank
    g = torch.from_numpy(np.concatenate([v.ravel() for v in new.get_globals()])).cuda()
    gs = [torch.empty_like(g) for _ in range(world)]
    dist.all_gather(gs, g)
    ok &= all(bool(torch.equal(gs[0], x)) for x in gs)
    print("rank %d K=%d %s: %s" % (rank, K, kind, "OK" if ok else "MISMATCH"), flush=True)
# the driver: hmmsgd_metaobs.VBHMM.infer with the sum over ranks inside the global-step kernel vs NCCL
from pysvihmm_b200 import hmmsgd_metaobs as H
from pysvihmm_b200.distributions import Gaussian
p = make_random_problem(seed=3, K=4, D=2, T_full=600, kind="niw_full", miss=0.1)
res = []
for peer in (True, False):
    objs = np.array([Gaussian(mu=e["mu"].copy(), sigma=e["sigma"].copy(), mu_0=pe["mu"], sigma_0=pe["sigma"],
                              kappa_0=pe["kappa"], nu_0=pe["nu"], kappa_mf=e["kappa"], nu_mf=e["nu"])
                     for e, pe in zip(p["emit"], p["prior_emit"])])
    hmm = H.VBHMM(p["obs"].copy(), np.ones(4), np.ones((4, 4)), objs, metaobs_half=8, mb_sz=3 * world, mask=p["mask"],
                  init_tran=p["var_tran"].copy(), maxit=4, seed=9, device=local, peer_allreduce=peer)
    hmm.infer()
    res.append((hmm.var_tran.copy(), np.array([g.mu_mf for g in hmm.var_emit]), hmm.elbo_vec.copy()))
okd = all(np.allclose(a, b, rtol=1e-9, atol=1e-10) for a, b in zip(res[0], res[1]))
print("rank %d VBHMM.infer peer vs nccl: %s" % (rank, "OK" if okd else "MISMATCH"), flush=True)
ok &= okd
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
