"""Data parallelism over meta-observations: the windows of one minibatch are independent given
the global parameters (the reference only *adds* their results, hmmsgd_metaobs.py:430-433), so
rank r takes windows r, r+G, r+2G, ... and ONE sum all-reduce of the packed statistics per global
step replaces the accumulation.  Every rank then applies the natural-gradient step redundantly on
bitwise-identical inputs, so replicas stay in lock-step without a broadcast."""
import numpy as np


def dist_or_none():
    """torch.distributed when initialised with world_size > 1, else None."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist
    except Exception:
        pass
    return None


def bind_to_gpu_numa_node(device_index):
    """Pin this process (and so the first-touch placement of the host buffers it allocates afterwards)
    to the CPUs of the NUMA node its GPU hangs off.  With one process per GPU and a host-resident series
    per rank, windows gathered over PCIe otherwise cross the inter-socket link for half of the ranks
    (measured at N = 8: gather 0.12 -> 0.20 ms per step).  Returns the node, or None when the topology
    files are not there (no-op)."""
    import os
    try:
        import torch
        props = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return None
        cpus = []
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.extend(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:       # noqa: BLE001  (sysfs layout, container restrictions: placement is an optimisation only)
        return None


def shard_starts(starts, rank, world):
    """This rank's share of the window start indices (strided, sizes differ by at most 1)."""
    return np.ascontiguousarray(np.asarray(starts, dtype=np.int64)[rank::world])


def broadcast_minibatch(starts, T, dist, device):
    """Rank 0's window starts and window length on every rank (one small broadcast per step).  The
    samplers draw from the process-local legacy numpy RNG; with different seeds (or seed=None) the
    ranks would otherwise shard DIFFERENT minibatches and the replicated globals would drift apart."""
    import torch
    n = torch.tensor([len(starts), int(T)], dtype=torch.int64)
    dev = device if dist.get_backend() == "nccl" else torch.device("cpu")
    n = n.to(dev)
    dist.broadcast(n, src=0)
    nb, T = int(n[0].item()), int(n[1].item())
    buf = torch.zeros(nb, dtype=torch.int64, device=dev)
    if dist.get_rank() == 0:
        buf.copy_(torch.as_tensor(np.asarray(starts, dtype=np.int64)))
    dist.broadcast(buf, src=0)
    return buf.cpu().numpy(), T


def allreduce_stats(stats, dist=None):
    """Sum the packed statistics tensor (include/svihmm.h layout) over ranks, in place."""
    dist = dist_or_none() if dist is None else dist
    if dist is not None:
        dist.all_reduce(stats)
    return stats


class PeerExchange(object):
    """One-shot all-reduce of the packed statistics over NVLink peer memory, fused into the
    global-step kernel (svihmm_global_update_peers) instead of a separate NCCL all-reduce.

    PyTorch is plumbing here: torch.distributed._symmetric_memory allocates one peer-accessible
    exchange area per rank (flags + receive slots) and hands every rank the device addresses of all
    of them; the exchange itself (P2P pushes, sequence flags, rank-ordered sums, update) is the
    hand-written kernel.  Usage per global step, on every rank:
        eng.estep(starts, T, stats=stats)            # (or estep_streamed) this rank's windows
        px.global_update(stats, lrate, bA, bE)       # replaces all_reduce + eng.global_update
    """

    def __init__(self, eng, dist):
        import ctypes as C

        import torch
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib as L
        self.eng, self.lib, self._L = eng, eng.lib, L
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        n = int(self.lib.svihmm_comm_buffer_len(eng._h))
        self.buf = symm_mem.empty(n, dtype=torch.float64, device=eng.device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, dist.group.WORLD)
        torch.cuda.synchronize(eng.device)
        dist.barrier()                                    # every rank's flags are zero before first use
        ptrs = (C.c_uint64 * self.world)(*[int(p) for p in self.hdl.buffer_ptrs])
        L.check(self.lib.svihmm_comm_attach(eng._h, self.rank, self.world, ptrs))

    def global_update(self, stats, lrate, bfact_A, bfact_E):
        import ctypes as C
        self._L.check(self.lib.svihmm_global_update_peers(self.eng._h, C.c_void_p(stats.data_ptr()), float(lrate),
                                                          float(bfact_A), float(bfact_E), self.eng._stream()))

    def reduced_stats(self, out=None):
        """The all-reduced statistics of the last step (device tensor)."""
        import ctypes as C

        import torch
        if out is None:
            out = torch.empty(self.eng.slen, dtype=torch.float64, device=self.eng.device)
        self._L.check(self.lib.svihmm_get_reduced_stats(self.eng._h, C.c_void_p(out.data_ptr()),
                                                        self._L.LOC_DEVICE, self.eng._stream()))
        return out
