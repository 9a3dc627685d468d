#!/usr/bin/env python
"""Benchmark of the SVI-HMM local E-step path: meta-observation E-steps/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|...]

One "step" = one global step of hmmsgd_metaobs.VBHMM.infer for a minibatch of B windows per GPU:
the batched E-step (emission log-likelihoods, forward, backward, marginals, sufficient statistics;
reference hmmsgd_metaobs.py:405-436), the all-reduce of the packed statistics when N > 1, and the
natural-gradient update (:1010-1069).  value = windows processed by all ranks / time (weak
scaling: B windows per GPU).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "meta-obs E-steps/sec (batched fwd-bwd, K states)"
CONFIGS = {
    # name: K, D, T, B per GPU, emission kind          (BASELINE.json configs[1] / configs[2])
    "c2": dict(K=16, D=8, T=512, B=256, kind="niw_diag",
               workload="configs[1]: K=16 diag-Gaussian, D=8, T=512, 256 meta-obs minibatch per GPU"),
    "c3": dict(K=64, D=32, T=1024, B=512, kind="niw_full",
               workload="configs[2]: K=64 full-cov Gaussian, D=32, T=1024, 512 meta-obs per GPU"),
    # configs[3]: the K x K step as a dense tensor-core contraction (tcgen05, bf16 messages); diagonal
    # emissions so that the float64 emission phase does not drown the step being measured
    # configs[4]: 4-component NIW mixtures per state (extension; per-phase kernels with K*C = 128 components)
    "c5": dict(K=32, D=16, T=2048, B=128, kind="niw_full", components=4,
               workload="configs[4]: K=32 states x 4-component full-cov GMM emissions, D=16, T=2048, "
                        "128 meta-obs per GPU"),
    "c4": dict(K=256, D=64, T=256, B=1024, kind="niw_diag", bf16_dense=True,
               workload="configs[3]: K=256 dense tensor-core K x K step (bf16), D=64 diag-Gaussian, T=256, "
                        "1024 meta-obs per GPU"),
}
T_FULL = 1 << 23          # rows of the synthetic series: 8.4M x D fp32 (268 MB at D=8) > 126 MB of L2


def synthetic_series(K, D, T_full, seed, sep=3.0):
    """Sticky-chain Gaussian HMM series, vectorised (state changes with prob 0.1 per step)."""
    rs = np.random.RandomState(seed)
    mus = sep * rs.randn(K, D)
    change = rs.rand(T_full) < 0.1
    seg = np.cumsum(change)
    seg_state = rs.randint(0, K, seg[-1] + 1)
    sts = seg_state[seg]
    obs = (mus[sts] + rs.standard_normal((T_full, D))).astype(np.float32)
    return obs, mus


def globals_for(K, D, kind, mus, seed):
    """var_tran = 1 + 5 rand; NIW per state around the true means."""
    rs = np.random.RandomState(seed + 1)
    var_tran = 1. + 5. * rs.rand(K, K)
    emit, prior = [], []
    for k in range(K):
        mu = mus[k] + 0.5 * rs.randn(D)
        if kind == "niw_full":
            nu = D + 3.
            emit.append(dict(mu=mu, sigma=(nu - D - 1) * np.eye(D), kappa=1., nu=nu))
            prior.append(dict(mu=np.zeros(D), sigma=np.eye(D), kappa=0.01, nu=D + 2.))
        else:
            emit.append(dict(mu=mu, sigma=2. * np.ones(D), kappa=np.ones(D), nu=4. * np.ones(D)))
            prior.append(dict(mu=np.zeros(D), sigma=np.ones(D), kappa=0.01 * np.ones(D), nu=3. * np.ones(D)))
    return var_tran, emit, prior


class ClockSampler(object):
    """SM clock and throttle reasons DURING the timed region, polled through NVML (the same
    counters as the nvidia-smi line of B200_PROFILING.md, at ~1 ms instead of >=50 ms period so
    that a timed region of a few milliseconds still gets samples)."""

    def __init__(self, gpu_index):
        self.idx, self.rows, self.stop_flag, self.thr, self.err = gpu_index, [], False, None, None

    def start(self):
        try:
            import pynvml as N
            N.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.idx]) if vis and vis.split(",")[self.idx].isdigit() else self.idx
            self.N, self.h = N, N.nvmlDeviceGetHandleByIndex(idx)
            self.max_sm = float(N.nvmlDeviceGetMaxClockInfo(self.h, N.NVML_CLOCK_SM))
            self.thr = threading.Thread(target=self._poll, daemon=True)
            self.thr.start()
        except Exception as e:      # noqa: BLE001
            self.err = repr(e)

    def _poll(self):
        N = self.N
        while not self.stop_flag:
            try:
                sm = N.nvmlDeviceGetClockInfo(self.h, N.NVML_CLOCK_SM)
                rs = N.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.rows.append((time.perf_counter(), float(sm), int(rs)))
            except Exception as e:  # noqa: BLE001
                self.err = repr(e)
                return
            time.sleep(0.0005)

    def stop(self, t0, t1):
        self.stop_flag = True
        if self.thr is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % self.err]}
        self.thr.join(1.0)
        N = self.N
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-1:]
        names = {"hw_slowdown": N.nvmlClocksEventReasonHwSlowdown,
                 "hw_thermal_slowdown": N.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": N.nvmlClocksEventReasonSwThermalSlowdown,
                 "sw_power_cap": N.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(k for k, bit in names.items() if any(r[2] & bit for r in rows))
        return {"sm_mhz": float(np.median([r[1] for r in rows])) if rows else None,
                "sm_max_mhz": self.max_sm, "reasons": reasons, "samples": len(rows)}


def nvlink_kib(gpu_index):
    """Cumulative NVLink payload counters of this GPU summed over its links (NVML field values, KiB) as
    [tx, rx], or a string saying why they could not be read: taken before and after the timed region at
    N > 1, the difference is what the fused peer exchange moved."""
    try:
        import pynvml as N
        N.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
        h = N.nvmlDeviceGetHandleByIndex(idx)
        ids = (N.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, N.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX)

        def val(x):
            t = int(x.valueType)        # NVML_VALUE_TYPE_*: 0 double, 1 uint, 2 ulong, 3 ulonglong, 4 slonglong
            return int({0: x.value.dVal, 1: x.value.uiVal, 2: x.value.ulVal, 3: x.value.ullVal,
                        4: x.value.sllVal}.get(t, x.value.ullVal))
        v = N.nvmlDeviceGetFieldValues(h, [(i, 0xFFFFFFFF) for i in ids])
        if all(int(x.nvmlReturn) == 0 for x in v):
            return [val(x) for x in v]
        rc_all = [int(x.nvmlReturn) for x in v]
        tot, seen = [0, 0], 0
        for link in range(18):          # per-link scope, summed
            v = N.nvmlDeviceGetFieldValues(h, [(i, link) for i in ids])
            if all(int(x.nvmlReturn) == 0 for x in v):
                tot[0] += val(v[0]); tot[1] += val(v[1]); seen += 1
        return tot if seen else "nvmlDeviceGetFieldValues: all-links rc %s, no per-link counter readable" % rc_all
    except Exception as e:       # noqa: BLE001
        return "nvml: %r" % (e,)


def ncu_traffic(kernel):
    """dram read+write bytes per launch of the dominant kernel from the committed ncu --set full capture
    (profiles/r2_c2_pipe_ncu_summary.txt), or None when there is no capture for this kernel/config."""
    if kernel != "fused":
        return None
    p = os.path.join(ROOT, "profiles", "r2_c2_pipe_ncu_summary.txt")
    try:
        tot, unit = 0.0, {"byte": 1., "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for line in open(p):
            if line.startswith("dram__bytes_read.sum") or line.startswith("dram__bytes_write.sum"):
                v = line.split("=")[1].split()
                tot += float(v[0]) * unit[v[1]]
        return tot or None
    except Exception:       # noqa: BLE001
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
# reference / CPU arm
# ------------------------------------------------------------------------------------------------
def _load_reference():
    """The reference's own classes (oracle/_ref = its sources patched mechanically to Python 3 by
    oracle/build_ref.py), or None."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref, "hmmsgd_metaobs.py")):
        return None
    import warnings
    warnings.filterwarnings("ignore")
    sys.path.insert(0, ref)
    try:
        import hmmsgd_metaobs as HS
        from pybasicbayes.distributions import Gaussian
        return HS, Gaussian
    except Exception:
        sys.path.remove(ref)
        return None


def cpu_estep_rate(cfg, budget_s, max_windows, seed=0):
    """E-steps/sec of the reference CPU path on a bounded sample of the workload: the loop body of
    hmmsgd_metaobs.py:405-436 (eig init, local_update, intermediate_pars, local_lower_bound) per
    window, float64, numpy with all the BLAS threads it takes.  The reference needs odd windows
    T = 2L+1, so it runs T-1 (=511 for c2) timesteps per E-step."""
    K, D, T = cfg["K"], cfg["D"], cfg["T"]
    Lh = (T - 1) // 2
    n = 200000
    obs32, mus = synthetic_series(K, D, n, seed)
    obs = obs32.astype(np.float64)
    var_tran, emit, prior = globals_for(K, D, cfg["kind"], mus, seed)
    starts = np.random.RandomState(seed + 2).randint(0, n - (2 * Lh + 1), max_windows)
    ref = _load_reference()
    done, t0 = 0, time.perf_counter()
    if ref is not None:
        HS, Gaussian = ref
        full = lambda v: np.diag(v) if np.ndim(v) == 1 else v          # diag config == full-cov with diagonal scale
        sc = lambda v: float(np.ravel(v)[0])
        objs = np.array([Gaussian(mu=e["mu"].copy(), sigma=full(e["sigma"]).copy(), mu_0=p["mu"],
                                  sigma_0=full(p["sigma"]), kappa_0=sc(p["kappa"]), nu_0=sc(p["nu"]) + D - 1,
                                  kappa_mf=sc(e["kappa"]), nu_mf=sc(e["nu"]) + D - 1)
                         for e, p in zip(emit, prior)])
        hmm = HS.VBHMM(obs, np.ones(K), np.ones((K, K)), objs, metaobs_half=Lh, mb_sz=max_windows,
                       init_tran=var_tran, maxit=1, seed=seed)
        t0 = time.perf_counter()
        for s in starts:
            mo = HS.MetaObs(int(s), int(s) + 2 * Lh)
            A_mean = hmm.var_tran / np.sum(hmm.var_tran, axis=1)[:, np.newaxis]     # :413-418
            ew, ev = np.linalg.eig(A_mean.T)
            hmm.var_init = np.abs(ev[:, np.argsort(ew)[::-1][0]])
            hmm.local_update(metaobs=mo)                                            # :422
            hmm.intermediate_pars(mo)                                               # :428
            hmm.local_lower_bound()                                                 # :436
            done += 1
            if time.perf_counter() - t0 > budget_s:
                break
        kind = "reference"
    else:
        from oracle import svihmm_oracle as O
        t0 = time.perf_counter()
        for s in starts:
            O.svi_minibatch_step(obs, None, [int(s)], 2 * Lh + 1, var_tran, emit, np.ones((K, K)), prior,
                                 0.5, Lh)
            done += 1
            if time.perf_counter() - t0 > budget_s:
                break
        kind = "port"
    dt = time.perf_counter() - t0
    return dict(value=done / dt, unit="E-steps/s", cores=os.cpu_count(), kind=kind,
                sample="%d windows of T=%d (reference needs odd 2L+1), K=%d D=%d, float64, %.1f s; "
                       "numpy/BLAS threads as the reference runs (single Python process)" % (
                           done, 2 * Lh + 1, K, D, dt)), done, dt


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = max(2, int(12.0 / max(args.steps, 1) * 36))      # ~12 s/step budget caps below
    t_all0 = time.perf_counter()
    tot, tt = 0, 0.0
    for i in range(args.warmup):
        cpu_estep_rate(cfg, 1.0, 8, seed=100 + i)
    budget = min(10.0, 150.0 / max(args.steps, 1))
    cb = None
    for i in range(args.steps):
        cb, done, dt = cpu_estep_rate(cfg, budget, per_step, seed=i)
        tot += done; tt += dt
    val = tot / tt
    cb["value"] = val
    cb["sample"] = "%d steps, each a sample of <=%d windows (%.0f s cap): " % (args.steps, per_step, budget) + cb["sample"]
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "E-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tt / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": cfg["workload"]},
            "cpu_baseline": cb,
            "e2e": {"value": val, "unit": "E-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": time.perf_counter() - t_all0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------

def quick_record(name, cfg, world, rank, local, dev, dist, B_per_gpu, seconds, scaling, steps=4, t_full=1 << 21):
    """A short (a few seconds) run of another BASELINE config beside the main one, so that the configs
    north_star ties to several GPUs are on the driver's record: same step as the main line (E-step over
    this rank's windows + sum over ranks inside the update + natural-gradient update), one
    svihmm_svi_run call per block of `steps` steps, CUDA events, median block, max over ranks.  The series
    is smaller than the main one (2M rows, still larger than L2)."""
    import torch
    from pysvihmm_b200 import _lib as L
    from pysvihmm_b200.engine import EStepEngine, pack_emit_dicts as pack_emit_np
    K, D, T, kind = cfg["K"], cfg["D"], cfg["T"], cfg["kind"]
    B = int(B_per_gpu)
    obs_host, mus = synthetic_series(K, D, t_full, seed=4242)
    C_mix = int(cfg.get("components", 1))
    var_tran, emit, prior = globals_for(K, D, kind, mus, seed=4242)
    if C_mix > 1:
        rsm = np.random.RandomState(7)
        emit = [dict(e, mu=e["mu"] + 0.7 * rsm.randn(D)) for e in emit for _ in range(C_mix)]
        prior = [p_ for p_ in prior for _ in range(C_mix)]
    eng = EStepEngine(K, D, kind, device=local, components=C_mix)
    eng.set_series(torch.from_numpy(obs_host).to(dev))
    del obs_host
    eng.set_prior(np.ones((K, K)), pack_emit_np(prior))
    if C_mix > 1:
        eng.set_mix_weights(2. * np.ones((K, C_mix)), np.ones((K, C_mix)))
    eng.set_globals(var_tran, pack_emit_np(emit))
    flags = L.WRAP | L.ADD_PRIOR | (L.BF16_DENSE if cfg.get("bf16_dense") else 0)
    Lh, S = T // 2, B * world
    bA = (t_full - 2 * Lh - 1) / (2. * Lh * S)
    bE = (t_full - 2 * Lh - 1) / ((2. * Lh + 1.) * S)
    px = None
    if world > 1:
        from pysvihmm_b200.sharding import PeerExchange
        px = PeerExchange(eng, dist)
    g = torch.Generator().manual_seed(77)
    pool = 4 * steps
    starts = torch.randint(0, t_full - T, (pool, world, B), generator=g)[:, rank].contiguous().to(dev)
    var_x = torch.empty((B, T, K), dtype=torch.float32, device=dev)
    stats = eng.new_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def block(i):
        i0 = (i % 4) * steps
        eng.svi_run(starts[i0:i0 + steps], T, 1.0, 0.7, i * steps, bA, bE, flags=flags, var_x=var_x, stats=stats,
                    peers=px is not None)
    block(0)
    sync()
    eng.set_profiling(True); eng.phase_ms()
    out, tot, i = [], 0.0, 1
    while True:
        sync()
        e0.record(); block(i); e1.record()
        sync()
        out.append(e0.elapsed_time(e1)); tot += out[-1]; i += 1
        go = torch.tensor([1.0 if (tot < 1e3 * seconds and len(out) < 200) else 0.0], device=dev)
        if world > 1:
            dist.broadcast(go, src=0)
        if go.item() < 0.5:
            break
    ph = eng.phase_ms()
    eng.set_profiling(False)
    tmax = torch.tensor([float(np.median(out))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item()) / steps
    rec = {"config": name, "workload": cfg["workload"], "B_per_gpu": B, "global_B": B * world, "scaling": scaling,
           "n_gpus": world, "ms_per_step": ms, "value": B * world / (ms * 1e-3), "unit": "E-steps/s",
           "blocks": len(out), "steps_per_block": steps,
           "phases_ms_per_step": {k: v[0] / (len(out) * steps) for k, v in ph.items()}}
    eng.close()
    del eng, var_x, starts
    torch.cuda.empty_cache()
    return rec


def xrank_check(eng, px, dist, starts_row, T, flags, bA, bE, dev):
    """Outside the timed region: the sum over ranks taken inside the update kernel (P2P pushes over
    NVLink) against an NCCL all-reduce of the same per-rank statistics, and bit-equality of the
    replicated globals across ranks after the update."""
    import torch
    stats = eng.new_stats()
    eng.estep(starts_row, T, flags=flags, want_var_x=False, stats=stats)
    ref = stats.clone()
    dist.all_reduce(ref)                                      # NCCL sum of the per-rank statistics
    px.global_update(stats, 0.1, bA, bE)
    red = px.reduced_stats()
    torch.cuda.synchronize()
    scale = float(ref.abs().max().item())
    err = float((red - ref).abs().max().item()) / max(scale, 1e-300)
    vt, vi, em = eng.get_globals()
    packed = torch.from_numpy(np.concatenate([vt.ravel(), vi.ravel(), em.ravel()])).to(dev).view(torch.int64)
    lo, hi = packed.clone(), packed.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool((lo == hi).all().item())
    ok = err < 1e-12 and same
    return {"xrank_check": "ok" if ok else "FAILED", "peer_sum_vs_nccl_rel_err": err, "globals_bitwise_equal_across_ranks": same}

def run_ours(args, cfg):
    # libraries (NCCL, symmetric memory) may print to stdout: keep fd 1 for the ONE JSON line
    json_fd = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    from pysvihmm_b200 import _lib as L
    from pysvihmm_b200.engine import EStepEngine, pack_emit_dicts as pack_emit_np
    from pysvihmm_b200.sharding import allreduce_stats

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_node = None
    if world > 1:
        from pysvihmm_b200.sharding import bind_to_gpu_numa_node
        numa_node = bind_to_gpu_numa_node(local)      # before the host series is allocated (first touch)
        dist.init_process_group("nccl", device_id=dev)
    K, D, T, B, kind = cfg["K"], cfg["D"], cfg["T"], cfg["B"], cfg["kind"]
    obs_host, mus = synthetic_series(K, D, T_FULL, seed=8675309)       # same series on every rank
    C_mix = int(cfg.get("components", 1))
    var_tran, emit, prior = globals_for(K, D, kind, mus, seed=8675309)
    if C_mix > 1:                       # component c of state k: the state's parameters with a jittered mean
        rsm = np.random.RandomState(7)
        emit = [dict(e, mu=e["mu"] + 0.7 * rsm.randn(D)) for e in emit for _ in range(C_mix)]
        prior = [p_ for p_ in prior for _ in range(C_mix)]
    eng = EStepEngine(K, D, kind, device=local, components=C_mix)
    eng.set_series(torch.from_numpy(obs_host).to(dev))
    eng.set_prior(np.ones((K, K)), pack_emit_np(prior))
    em0 = pack_emit_np(emit)
    if C_mix > 1:
        eng.set_mix_weights(2. * np.ones((K, C_mix)), np.ones((K, C_mix)))
    eng.set_globals(var_tran, em0)
    flags = L.WRAP | L.ADD_PRIOR | (L.BF16_DENSE if cfg.get("bf16_dense") else 0)
    if args.B:
        B = int(args.B)
    Lh, S = T // 2, B * world
    bA = (T_FULL - 2 * Lh - 1) / (2. * Lh * S)
    bE = (T_FULL - 2 * Lh - 1) / ((2. * Lh + 1.) * S)
    if args.B:
        B = int(args.B)
    # every step draws fresh windows (as metaobs_unif does): rank r takes its own B of the N*B; a pool of
    # `nst` minibatches (>= 1.3 GB of distinct window rows at c2) is cycled through by the repeated blocks
    nst = max(args.warmup + args.steps, min(64, max(8, (1 << 22) // max(B, 1))))
    g = torch.Generator().manual_seed(1234)
    all_starts = torch.randint(0, T_FULL - T, (nst, world, B), generator=g)[:, rank].contiguous()
    starts_dev = all_starts.to(dev)
    var_x = torch.empty((B, T, K), dtype=torch.float32, device=dev)
    stats = eng.new_stats()

    # N > 1: the sum over ranks of the statistics is taken inside the global-step kernel over NVLink
    # peer memory (sharding.PeerExchange); --nccl keeps the plain NCCL all-reduce for comparison
    px = None
    if world > 1 and not args.nccl:
        from pysvihmm_b200.sharding import PeerExchange
        try:
            px = PeerExchange(eng, dist)
            okflag = torch.ones(1, device=dev)
        except Exception as e:          # noqa: BLE001  (no peer access / symmetric memory on this box)
            sys.stderr.write("rank %d: peer exchange unavailable (%r), using the NCCL all-reduce\n" % (rank, e))
            px, okflag = None, torch.zeros(1, device=dev)
        dist.all_reduce(okflag, op=dist.ReduceOp.MIN)          # all ranks must agree on the path
        if okflag.item() < 1:
            px = None

    def step(i, it):
        i = i % nst
        eng.estep(starts_dev[i], T, flags=flags, var_x=var_x, stats=stats)
        if px is not None:
            px.global_update(stats, (it + 1.) ** -0.7, bA, bE)
            return
        if world > 1:
            allreduce_stats(stats, dist)
        eng.global_update(stats, (it + 1.) ** -0.7, bA, bE)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i, i)
    sync()
    xr = None
    if px is not None:
        xr = xrank_check(eng, px, dist, starts_dev[0], T, flags, bA, bE, dev)
        eng.set_globals(var_tran, em0)
        sync()
    # ---- timed region 1: inputs resident in HBM ------------------------------------------------
    # (no per-phase events inside the timed region: an event record between two kernels would break the
    # programmatic-dependent-launch chain of svihmm_svi_run; the per-kernel durations for the roofline are
    # taken by the same CUDA-event mechanism in a second pass over the same blocks right after it)
    l0 = eng.launch_count()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.15)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    use_run = (world == 1 or px is not None) and not args.py_loop
    assert nst >= args.steps

    def run_block(it):
        """One block of args.steps global steps through ONE C call (svihmm_svi_run): the pool of
        minibatches is walked in consecutive slices."""
        i0 = (it // args.steps) % (nst // args.steps) * args.steps
        eng.svi_run(starts_dev[i0:i0 + args.steps], T, 1.0, 0.7, it, bA, bE, flags=flags, var_x=var_x, stats=stats,
                    peers=px is not None)

    def timed_blocks(fn, min_seconds, max_blocks=100000):
        """Blocks of EXACTLY args.steps steps, each bracketed by barrier + synchronize on both sides and
        timed with CUDA events on the launching stream, repeated until the summed timed region reaches
        min_seconds (so that clock / utilisation samplers see the load).  Every rank runs the same number
        of blocks (rank 0 decides).  Returns the per-block milliseconds of this rank."""
        out, tot, it = [], 0.0, args.warmup
        while True:
            sync()
            e0.record()
            if fn is step and use_run:
                run_block(it)
                it += args.steps
            else:
                for _ in range(args.steps):
                    fn(it, it)
                    it += 1
            e1.record()
            sync()
            out.append(e0.elapsed_time(e1))
            tot += out[-1]
            go = torch.tensor([1.0 if (tot < 1e3 * min_seconds and len(out) < max_blocks) else 0.0], device=dev)
            if world > 1:
                dist.broadcast(go, src=0)
            if go.item() < 0.5:
                return out

    nvl0 = nvlink_kib(local) if (world > 1 and rank == 0) else None
    tw0 = time.perf_counter()
    blocks = timed_blocks(step, args.min_seconds)
    tw1 = time.perf_counter()
    nvl1 = nvlink_kib(local) if (world > 1 and rank == 0) else None
    launches = (eng.launch_count() - l0) // len(blocks)
    clk = clocks.stop(tw0, tw1) if rank == 0 else None
    eng.set_profiling(True)
    eng.phase_ms()
    nprof = min(len(blocks), 50)
    it_p = args.warmup
    for _ in range(nprof):
        if use_run:
            run_block(it_p); it_p += args.steps
        else:
            for _ in range(args.steps):
                step(it_p, it_p); it_p += 1
    sync()
    phases = eng.phase_ms()
    phases = {k: (v[0] / nprof, v[1] // nprof) for k, v in phases.items()}
    eng.set_profiling(False)
    # median block of this rank, then the max over ranks
    tmax = torch.tensor([float(np.median(blocks))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    value = world * B * args.steps / (ms * 1e-3)
    timed_region_s = float(np.sum(blocks)) * 1e-3

    # ---- B sweep (N = 1): where the E-step saturates (the benchmark's B = 256 is latency-bound) ----
    b_sweep = None
    sweep_list = [int(x) for x in args.sweep.split(",") if x] if args.sweep else []
    if world == 1 and sweep_list:
        b_sweep = []
        alg_bytes_ = T * (D * 4 + K * 4)
        peak_, _ = measured_peaks()
        for Bs in sweep_list:
            g2 = torch.Generator().manual_seed(99 + Bs)
            npool = max(4, min(32, (1 << 21) // Bs))
            st2 = torch.randint(0, T_FULL - T, (npool, Bs), generator=g2).to(dev)
            vx2 = torch.empty((Bs, T, K), dtype=torch.float32, device=dev)
            bA2 = (T_FULL - 2 * Lh - 1) / (2. * Lh * Bs); bE2 = (T_FULL - 2 * Lh - 1) / ((2. * Lh + 1.) * Bs)

            def step2(i, it):
                eng.estep(st2[i % npool], T, flags=flags, var_x=vx2, stats=stats)
                eng.global_update(stats, (it + 1.) ** -0.7, bA2, bE2)
            for i in range(3):
                step2(i, i)
            sync()
            eng.set_profiling(True); eng.phase_ms()
            nrep, tot, it2 = 0, 0.0, 3
            while tot < 1e3 * args.sweep_seconds and nrep < 2000:
                e0.record()
                step2(it2, it2)
                e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1); nrep += 1; it2 += 1
            ph = eng.phase_ms(); eng.set_profiling(False)
            est_ms = sum(v[0] for k_, v in ph.items() if k_ != "update") / nrep
            b_sweep.append({"B": Bs, "ms_per_step": tot / nrep, "estep_ms": est_ms, "steps": nrep,
                            "value": Bs / (tot / nrep * 1e-3),
                            "estep_frac_of_hbm_peak": Bs * alg_bytes_ / (est_ms * 1e-3) / 1e9 / peak_,
                            "phases_ms": {k_: v[0] / nrep for k_, v in ph.items()}})
            del vx2, st2
        eng.set_globals(var_tran, em0)

    # ---- timed region 2: end to end through the host-buffer call --------------------------------
    if C_mix > 1:
        eng.set_mix_weights(2. * np.ones((K, C_mix)), np.ones((K, C_mix)))
    eng.set_globals(var_tran, em0)
    eng.set_series_streamed(obs_host)
    stats_h = np.empty(eng.slen)
    starts_h = all_starts.numpy()
    stats_d = eng.new_stats()
    stats_pin = torch.empty(eng.slen, dtype=torch.float64).pin_memory()

    def step_host(i, it):
        i = i % nst
        # the call a user makes per global step, HOST buffers in and out: this step's windows come from
        # the host series (H2D inside the timed region; the NEXT step's windows are announced so that
        # their gather overlaps this step's compute), the minibatch statistics are read back (D2H)
        nxt = starts_h[(i + 1) % nst]
        eng.prefetch_windows(starts_h[(i + 2) % nst], T)    # two minibatches ahead: the host link never idles
        lr = (it + 1.) ** -0.7
        if world == 1:
            eng.svi_step_host(starts_h[i], T, lr, bA, bE, next_starts=nxt, flags=flags, stats_out=stats_h)
        elif px is not None:
            eng.estep_streamed(starts_h[i], T, next_starts=nxt, flags=flags, stats=stats_d)
            px.global_update(stats_d, lr, bA, bE)
            stats_pin.copy_(px.reduced_stats(stats_d), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        else:
            eng.estep_streamed(starts_h[i], T, next_starts=nxt, flags=flags, stats=stats_d)
            allreduce_stats(stats_d, dist)
            eng.global_update(stats_d, lr, bA, bE)
            stats_pin.copy_(stats_d, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    for i in range(args.warmup):
        step_host(i, i)
    blocks2 = timed_blocks(step_host, min(args.min_seconds, 1.0))
    # untimed repeat with per-phase events: where the end-to-end step spends its device time
    eng.set_profiling(True)
    eng.phase_ms()
    for i in range(args.warmup, args.warmup + 20):
        step_host(i, i)
    sync()
    phases2 = eng.phase_ms()
    eng.set_profiling(False)
    tmax = torch.tensor([float(np.median(blocks2))], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms2 = float(tmax.item())
    e2e = world * B * args.steps / (ms2 * 1e-3)

    # ---- the other BASELINE configs, briefly (driver-visible records for the multi-GPU configs) ----
    extras = []
    slen = eng.slen
    if args.extras and args.config == "c2" and not args.B:
        eng.close()
        del eng, var_x, starts_dev
        torch.cuda.empty_cache()
        try:
            # configs[2]: 4096 windows in all, sharded over the ranks (strong scaling)
            extras.append(quick_record("c3", CONFIGS["c3"], world, rank, local, dev, dist, 4096 // world, 1.0, "strong"))
            # configs[3]: 1024 windows per GPU (weak scaling sweep)
            extras.append(quick_record("c4", CONFIGS["c4"], world, rank, local, dev, dist, 1024, 1.0, "weak"))
            # configs[4]: 512 windows in all (strong scaling)
            extras.append(quick_record("c5", CONFIGS["c5"], world, rank, local, dev, dist, max(512 // world, 1), 1.0, "strong"))
        except Exception as e:          # noqa: BLE001
            extras.append({"error": repr(e)})
        eng = None

    if rank == 0:
        peak, peak_src = measured_peaks()
        alg_bytes = T * (D * 4 + K * 4)                      # read window once, write var_x once
        # the batched forward-backward = all kernels of the E-step (one fused kernel, or emit + chain +
        # post of the batched path): their summed CUDA-event durations per step, update excluded
        est = {k: v for k, v in phases.items() if k not in ("update", "gather")}
        dom = ("+".join(sorted(est)) if est else "none", (sum(v[0] for v in est.values()), args.steps))
        dom_ms = dom[1][0] / max(dom[1][1], 1)
        ach = B * alg_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        step_ach = B * alg_bytes / (ms / args.steps * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "E-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "repeats": len(blocks), "timed_region_s": timed_region_s,
            "timing": "median over `repeats` blocks of exactly `steps` steps (each block: barrier + synchronize, "
                      "CUDA events on the launching stream), max over ranks",
            "scaling": "weak", "vs_baseline": None,
            "dtype": ("bf16 messages / f32 accumulate (tcgen05) recursions, " if cfg.get("bf16_dense") else "f32 recursions, ")
                     + "f64 emission log-lik + statistics",
            "data": "synthetic",
            "config": {"workload": cfg["workload"], "K": K, "D": D, "T": T, "B_per_gpu": B,
                       "series": "%d x %d fp32 (%.0f MB) resident in HBM, larger than L2; fresh random "
                                 "windows every step (no L2 flush needed)" % (T_FULL, D, T_FULL * D * 4 / 1e6),
                       "driver": ("svihmm_svi_run: one C call enqueues the %d steps of a block" % args.steps) if use_run
                                 else "Python loop: svihmm_estep + svihmm_global_update per step",
                       "step": "E-step (B windows) + %sglobal natural-gradient update" % (
                           ("" if world == 1 else "NCCL all-reduce of packed statistics + " if px is None else
                            "sum over ranks by P2P pushes over NVLink inside the ")),
                       "var_x_written": True, "rank0_numa_node": numa_node},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": ncu_traffic(dom[0]) if args.config == "c2" else None, "kernel": dom[0], "kernel_ms": dom_ms, "peak_source": peak_src,
                         "algorithmic_bytes_per_estep": alg_bytes,
                         "whole_step_achieved": step_ach, "whole_step_frac": step_ach / peak,
                         "phases_ms_per_step": {k: v[0] / args.steps for k, v in phases.items()}},
            "e2e": {"value": e2e, "unit": "E-steps/s", "ms_per_step": ms2 / args.steps,
                    "h2d_bytes_per_step": B * T * D * 4 + B * 8, "d2h_bytes_per_step": slen * 8,
                    "phases_ms_per_call": {k: v[0] / max(v[1], 1) for k, v in phases2.items()},
                    "call": "svihmm_svi_step_host (one C-ABI call per global step: host windows in, statistics "
                            "out; the next step's windows are gathered over PCIe while this step computes)"
                            if world == 1 else ("svihmm_estep_streamed + svihmm_global_update_peers (sum over ranks by P2P "
                                                "pushes over NVLink inside the update kernel) + D2H" if px is not None else
                                                "svihmm_estep_streamed + NCCL all-reduce + svihmm_global_update + D2H"),
                    "repeats": len(blocks2)},
            "gpu_launches": int(launches), "clocks": clk,
        }
        if b_sweep is not None:
            line["b_sweep"] = b_sweep
        if xr is not None:
            line.update(xr)
        if isinstance(nvl0, str) or isinstance(nvl1, str):
            line["nvlink"] = {"unavailable": nvl0 if isinstance(nvl0, str) else nvl1}
        elif nvl0 is not None and nvl1 is not None:
            # rank 0's NVLink payload counters over the timed region (NVML, all links): what the peer exchange moved
            nsteps = len(blocks) * args.steps
            line["nvlink"] = {"source": "NVML NVLINK_THROUGHPUT_DATA_TX/RX of rank 0's GPU, difference over the timed region",
                              "tx_bytes_per_step": (nvl1[0] - nvl0[0]) * 1024.0 / nsteps,
                              "rx_bytes_per_step": (nvl1[1] - nvl0[1]) * 1024.0 / nsteps,
                              "statistics_bytes_per_rank": slen * 8, "steps_counted": nsteps}
        if extras:
            line["extra_configs"] = extras
        if cfg.get("bf16_dense"):
            # the dense K x K work (forward + backward matvecs + transition statistic =
            # 6 K^2 T flop per E-step) against the measured bf16 tensor peak, over the phases that hold it
            try:
                pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
                tpeak, tsrc = float(pk["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
            except Exception:
                tpeak, tsrc = 1357.7, "fallback (SURVEY section 8d bf16_tflops_sustained)"
            ph = line["roofline"]["phases_ms_per_step"]
            dense_ms = ph.get("forward", 0.0) + ph.get("stats", 0.0)
            if dense_ms > 0:
                tf = B * 6.0 * K * K * T / (dense_ms * 1e-3) / 1e12
                line["roofline"]["tensor"] = {"bound": "tensor", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s",
                                              "frac": tf / tpeak, "flops_fb_per_estep": 6 * K * K * T,
                                              "phases": "forward (both chains) + stats", "peak_source": tsrc}
        if world == 1 and not args.no_cpu and C_mix == 1:     # the reference has no mixture-emission E-step
            cb, _, _ = cpu_estep_rate(cfg, 12.0, 400)
            line["cpu_baseline"] = cb
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--nccl", action="store_true", help="N > 1: plain NCCL all-reduce instead of the fused peer sum")
    ap.add_argument("--min-seconds", type=float, default=2.0,
                    help="repeat the timed block of --steps steps until the timed region is at least this long")
    ap.add_argument("--no-extras", dest="extras", action="store_false",
                    help="skip the short records of configs c3 / c4 / c5 appended to the c2 line")
    ap.add_argument("--py-loop", action="store_true",
                    help="drive every global step from Python (svihmm_estep + svihmm_global_update per step) "
                         "instead of one svihmm_svi_run call per block")
    ap.add_argument("--sweep", default=None,
                    help="comma-separated windows-per-GPU values for the saturation sweep (N = 1); default: "
                         "1024,4096,16384 for c2, none otherwise; '' disables")
    ap.add_argument("--sweep-seconds", type=float, default=0.25)
    ap.add_argument("--B", type=int, default=0, help="override the windows per GPU of the config (B sweep)")
    args = ap.parse_args()
    if args.sweep is None:
        args.sweep = "1024,4096,16384" if (args.config == "c2" and not args.B) else ""
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()
