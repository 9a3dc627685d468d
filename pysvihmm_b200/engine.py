"""EStepEngine: thin Python holder of PyTorch tensors around the C ABI of libsvihmm.so.

PyTorch is used for device memory, streams and (by the callers) torch.distributed only; all
arithmetic of the path happens in the hand-written CUDA kernels behind include/svihmm.h.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib as L

_KINDS = {"niw_full": L.EMIT_NIW_FULL, "niw_diag": L.EMIT_NIW_DIAG, "categorical": L.EMIT_CATEGORICAL}


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return C.c_void_p(t.data_ptr())
    if isinstance(t, np.ndarray):
        return C.c_void_p(t.ctypes.data)
    raise TypeError(type(t))


def _f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def pack_emit_dicts(emit):
    """list of dict(mu, sigma, kappa, nu) -> (K, plen) float64 in the layout of include/svihmm.h
    (sigma 2-D = full NIW, 1-D = diagonal)."""
    rows = []
    for e in emit:
        if "alpha" in e:                     # categorical: Dirichlet parameters
            rows.append(np.asarray(e["alpha"], dtype=np.float64).ravel())
            continue
        mu = np.asarray(e["mu"], dtype=np.float64).ravel()
        D = mu.size
        sg = np.asarray(e["sigma"], dtype=np.float64)
        if sg.ndim == 2:
            rows.append(np.concatenate([mu, sg.ravel(), [float(e["kappa"])], [float(e["nu"])]]))
        else:
            rows.append(np.concatenate([mu, sg, np.broadcast_to(np.asarray(e["kappa"], float), (D,)),
                                        np.broadcast_to(np.asarray(e["nu"], float), (D,))]))
    return np.array(rows)


class EStepEngine(object):
    """One engine = one svihmm_ctx on one GPU.

    Replaces, for a whole minibatch at once, the reference's local_update + intermediate_pars
    (hmmsgd_metaobs.py:405-436) and global_update (:1010-1069) / hmmbatchcd.global_update
    (hmmbatchcd.py:172-189).
    """

    def __init__(self, K, D, emission="niw_full", device=None, components=1):
        if not torch.cuda.is_available():
            raise L.SvihmmError("pysvihmm_b200 needs a CUDA device (no CPU fallback)")
        self.lib = L.load()
        self.K, self.D = int(K), int(D)
        self.emission = emission
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device) \
            if not isinstance(device, torch.device) else device
        h = C.c_void_p()
        self.C = int(components)                 # mixture components per state (EXTENSION, config 5)
        self.KE = self.K * self.C                # emission components: rows of the emission arrays
        if self.C > 1:
            L.check(self.lib.svihmm_create_mix(C.byref(h), self.device.index, self.K, self.D, _KINDS[emission],
                                               self.C))
        else:
            L.check(self.lib.svihmm_create(C.byref(h), self.device.index, self.K, self.D, _KINDS[emission]))
        self._h = h
        self.plen = int(self.lib.svihmm_emit_param_len(h))
        self.slen = int(self.lib.svihmm_stats_len(h))
        self.DD = {"niw_full": self.D * self.D, "niw_diag": self.D, "categorical": 0}[emission]
        self.OD = 1 if emission == "categorical" else self.D     # columns of the observation series
        self._keep = {}          # borrowed tensors / host arrays the C side points into
        self.T_full = None

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None):
            self.lib.svihmm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ------------------------------------------------------------------ data
    def set_series(self, obs, mask=None, dtype=None):
        """obs: (T_full, D) torch CUDA tensor (borrowed) or numpy array (copied to HBM).
        mask: (T_full,) bool, True = missing (hmmbase.py:60-65).  dtype: 'f32'/'f64' to convert."""
        if isinstance(obs, np.ndarray):
            obs = torch.from_numpy(np.ascontiguousarray(obs.reshape(obs.shape[0], -1)))
        if dtype is not None:
            obs = obs.to(torch.float32 if dtype == "f32" else torch.float64)
        if obs.dtype not in (torch.float32, torch.float64):
            obs = obs.to(torch.float64)
        obs = obs.to(self.device).contiguous().reshape(obs.shape[0], -1)
        if obs.shape[1] != self.OD:
            raise ValueError("obs has %d columns, engine expects %d" % (obs.shape[1], self.OD))
        m = None
        if mask is not None:
            m = torch.as_tensor(np.asarray(mask, dtype=np.uint8) if not isinstance(mask, torch.Tensor)
                                else mask.to(torch.uint8)).to(self.device).contiguous()
        self._keep["obs"], self._keep["mask"] = obs, m
        self.T_full = int(obs.shape[0])
        L.check(self.lib.svihmm_set_series(self._h, _ptr(obs), self.T_full,
                                           L.F32 if obs.dtype == torch.float32 else L.F64,
                                           _ptr(m), L.LOC_DEVICE, self._stream()))

    def set_series_streamed(self, obs_host, mask_host=None):
        """Host-resident series (numpy, C-contiguous f32/f64): windows are gathered per step by
        estep_host (gen_synthetic.read_data_mmap feeder, gen_synthetic.py:188-191)."""
        obs_host = np.ascontiguousarray(obs_host)
        if obs_host.dtype not in (np.float32, np.float64):
            obs_host = obs_host.astype(np.float64)
        obs_host = obs_host.reshape(obs_host.shape[0], -1)
        m = None if mask_host is None else np.ascontiguousarray(np.asarray(mask_host, dtype=np.uint8))
        self._keep["hobs"], self._keep["hmask"] = obs_host, m
        self.hT_full = int(obs_host.shape[0])
        L.check(self.lib.svihmm_set_series_streamed(
            self._h, _ptr(obs_host), self.hT_full, L.F32 if obs_host.dtype == np.float32 else L.F64,
            _ptr(m)))

    def set_series_memmap(self, fname, T, D=None, mask_host=None, page_lock=True):
        """The on-disk series of gen_synthetic.generate_data_mmap / read_data_mmap (gen_synthetic.py:158-191:
        a float64 memmap of shape (T, D)) as the streamed series: windows are gathered per step by
        prefetch_windows / estep_streamed / svi_step_host, so the series may be larger than HBM.
        page_lock=False: never page-lock the mapping (series larger than host memory): the CPU gathers
        each minibatch's windows into pinned staging, the OS pages the file in on demand."""
        mm = np.memmap(fname, dtype=np.float64, mode="r", shape=(int(T), int(self.OD if D is None else D)))
        if not page_lock:
            self.set_tuning(L.TUNE_NO_HOSTREG, 1)
        self.set_series_streamed(mm, mask_host)
        return mm

    # ------------------------------------------------------------------ parameters
    def set_prior(self, prior_tran, prior_emit, prior_init=None):
        pt, pe = _f64(prior_tran), _f64(prior_emit)
        assert pt.shape == (self.K, self.K) and pe.size == self.KE * self.plen
        pi = None if prior_init is None else _f64(prior_init)
        L.check(self.lib.svihmm_set_prior(self._h, _ptr(pt), _ptr(pi), _ptr(pe), L.LOC_HOST, self._stream()))

    def set_globals(self, var_tran, emit, var_init=None):
        vt, em = _f64(var_tran), _f64(emit)
        assert vt.shape == (self.K, self.K) and em.size == self.KE * self.plen
        vi = None if var_init is None else _f64(var_init)
        L.check(self.lib.svihmm_set_globals(self._h, _ptr(vt), _ptr(vi), _ptr(em), L.LOC_HOST, self._stream()))

    def check(self):
        """svihmm_check: raises SvihmmError if a global step lost an emission scale to cancellation."""
        L.check(self.lib.svihmm_check(self._h, self._stream()))

    def get_globals(self):
        self.check()
        vt = np.empty((self.K, self.K)); vi = np.empty(self.K); em = np.empty((self.KE, self.plen))
        L.check(self.lib.svihmm_get_globals(self._h, _ptr(vt), _ptr(vi), _ptr(em), L.LOC_HOST, self._stream()))
        return vt, vi, em

    def set_mix_weights(self, omega, omega_prior=None):
        """Dirichlet parameters (K, C) of the mixture weights and, on the first call, their prior.
        Call before set_globals."""
        om = _f64(omega)
        assert om.size == self.KE
        op = None if omega_prior is None else _f64(omega_prior)
        L.check(self.lib.svihmm_set_mix_weights(self._h, _ptr(om), _ptr(op), L.LOC_HOST, self._stream()))

    def get_mix_weights(self):
        om = np.empty((self.K, self.C))
        L.check(self.lib.svihmm_get_mix_weights(self._h, _ptr(om), L.LOC_HOST, self._stream()))
        return om

    # ------------------------------------------------------------------ the hot path
    def new_stats(self):
        return torch.empty(self.slen, dtype=torch.float64, device=self.device)

    def set_var_init(self, var_init=None):
        """Explicit initial-state Dirichlet parameter (K,), or None = stationary vector of the current
        transition parameters (hmmsgd_metaobs.py:413-418)."""
        vi = None if var_init is None else _f64(var_init)
        L.check(self.lib.svihmm_set_var_init(self._h, _ptr(vi), L.LOC_HOST, self._stream()))

    def estep(self, starts, T, flags=0, var_x=None, stats=None, want_var_x=True, keep_locals=False, trim=0):
        """E-step over windows obs[starts[b]:starts[b]+T] of the resident series.
        Returns (var_x (B,T,K) float32 CUDA tensor or None, stats float64 CUDA tensor).
        keep_locals: keep the lliks/alpha/scale tables for get_locals (unfused kernels)."""
        if keep_locals:
            flags = int(flags) | L.KEEP_LOCALS
        if not (isinstance(starts, torch.Tensor) and starts.is_cuda):
            # host-side window starts are validated here (the kernels index the resident series
            # unchecked); device-resident starts are the caller's responsibility (no sync on the hot path)
            sh = np.asarray(starts.numpy() if isinstance(starts, torch.Tensor) else starts, dtype=np.int64).ravel()
            if self.T_full is not None and sh.size and (sh.min() < 0 or sh.max() + int(T) > self.T_full):
                raise L.SvihmmError("window [%d, +%d) outside the series of length %d" % (
                    int(sh.min() if sh.min() < 0 else sh.max()), int(T), self.T_full))
            starts = torch.from_numpy(np.ascontiguousarray(sh))
        starts = starts.to(device=self.device, dtype=torch.int64).contiguous()
        B = int(starts.numel())
        if var_x is None and want_var_x:
            var_x = torch.empty((B, T, self.K), dtype=torch.float32, device=self.device)
        if stats is None:
            stats = self.new_stats()
        if trim:          # buffered windows: statistics from the inner T - 2*trim rows only
            L.check(self.lib.svihmm_estep_buffered(self._h, _ptr(starts), B, int(T), int(trim), _ptr(var_x),
                                                   _ptr(stats), int(flags), self._stream()))
        else:
            L.check(self.lib.svihmm_estep(self._h, _ptr(starts), B, int(T), _ptr(var_x), _ptr(stats),
                                          int(flags), self._stream()))
        self._keep["starts"] = starts
        return var_x, stats

    def estep_host(self, starts, T, flags=0, want_var_x=False, stats_out=None, var_x_out=None):
        """Reference-facing call with HOST buffers: gathers the windows from the streamed host
        series host->device, runs the E-step, copies stats (and var_x) back, synchronises."""
        starts = np.ascontiguousarray(np.asarray(starts, dtype=np.int64))
        B = int(starts.size)
        stats = np.empty(self.slen) if stats_out is None else stats_out
        vx = var_x_out
        if vx is None and want_var_x:
            vx = np.empty((B, T, self.K), dtype=np.float32)
        L.check(self.lib.svihmm_estep_host(self._h, _ptr(starts), B, int(T), _ptr(vx), _ptr(stats),
                                           int(flags), self._stream()))
        return vx, stats

    def prefetch_windows(self, starts, T):
        """Announce an upcoming minibatch: its windows start moving host->device on the engine's
        copy stream now (ring of 3 staging slots); estep_streamed / svi_step_host with the same
        starts later use the staged copy."""
        starts = np.ascontiguousarray(np.asarray(starts, dtype=np.int64))
        L.check(self.lib.svihmm_prefetch_windows(self._h, _ptr(starts), int(starts.size), int(T)))

    def estep_streamed(self, starts, T, next_starts=None, flags=0, stats=None, var_x=None):
        """Enqueue-only E-step over windows of the HOST-resident series (set_series_streamed):
        the windows are gathered host->device on the current stream, or taken from the staging
        buffer if `starts` was announced as `next_starts` of the previous call, in which case the
        gather overlapped the previous step.  Statistics stay on the device (all-reduce /
        global_update follow on the same stream).  Returns the stats tensor."""
        starts = np.ascontiguousarray(np.asarray(starts, dtype=np.int64))
        nxt = None if next_starts is None else np.ascontiguousarray(np.asarray(next_starts, dtype=np.int64))
        if nxt is not None and nxt.size != starts.size:
            raise ValueError("next_starts must have as many windows as starts")
        if stats is None:
            stats = self.new_stats()
        L.check(self.lib.svihmm_estep_streamed(self._h, _ptr(starts), int(starts.size), int(T), _ptr(nxt),
                                               _ptr(var_x), _ptr(stats), int(flags), self._stream()))
        return stats

    def svi_step_host(self, starts, T, lrate, bfact_A, bfact_E, next_starts=None, flags=0, stats_out=None):
        """One global step of hmmsgd_metaobs.VBHMM.infer (:396-439) with HOST buffers: windows in
        (gathered from the streamed host series), minibatch statistics out (numpy float64), the
        natural-gradient update applied to the device-resident globals.  Synchronises."""
        starts = np.ascontiguousarray(np.asarray(starts, dtype=np.int64))
        nxt = None if next_starts is None else np.ascontiguousarray(np.asarray(next_starts, dtype=np.int64))
        if nxt is not None and nxt.size != starts.size:
            raise ValueError("next_starts must have as many windows as starts")
        stats = np.empty(self.slen) if stats_out is None else stats_out
        L.check(self.lib.svihmm_svi_step_host(self._h, _ptr(starts), int(starts.size), int(T), _ptr(nxt),
                                              _ptr(stats), int(flags), float(lrate), float(bfact_A),
                                              float(bfact_E), self._stream()))
        return stats

    def global_update(self, stats, lrate, bfact_A, bfact_E):
        """hmmsgd_metaobs.py:1010-1069 on the device-resident globals."""
        L.check(self.lib.svihmm_global_update(self._h, _ptr(stats), float(lrate), float(bfact_A),
                                              float(bfact_E), self._stream()))

    def set_adagrad(self, on=True):
        """AdaGrad-like transition step (hmmsgd_metaobs.py:1036-1040); on=True resets ada_G to ones."""
        L.check(self.lib.svihmm_set_adagrad(self._h, int(bool(on)), self._stream()))

    def batch_update(self, stats):
        """hmmbatchcd.py:172-189 on the device-resident globals."""
        L.check(self.lib.svihmm_batch_update(self._h, _ptr(stats), self._stream()))

    def batchsgd_update(self, stats, lrate):
        """hmmbatchsgd.py:202-259 on the device-resident globals."""
        L.check(self.lib.svihmm_batchsgd_update(self._h, _ptr(stats), float(lrate), self._stream()))

    def get_locals(self, B, T, tables=True):
        """Per-window [logZ, Q4 bound]; with tables=True also lliks/alpha/mx/cs (needs the last
        estep to have run with keep_locals=True)."""
        lz = np.empty((B, 2))
        if not tables:
            L.check(self.lib.svihmm_get_locals(self._h, None, None, None, None, _ptr(lz), L.LOC_HOST,
                                               self._stream()))
            return dict(logZ=lz[:, 0], lb_q4=lz[:, 1])
        ll = np.empty((B, T, self.K)); al = np.empty((B, T, self.K), dtype=np.float32)
        mx = np.empty((B, T)); cs = np.empty((B, T), dtype=np.float32)
        L.check(self.lib.svihmm_get_locals(self._h, _ptr(ll), _ptr(al), _ptr(mx), _ptr(cs), _ptr(lz),
                                           L.LOC_HOST, self._stream()))
        be = np.empty((B, T, self.K), dtype=np.float32); sb = np.empty((B, T), dtype=np.float32)
        L.check(self.lib.svihmm_get_locals_beta(self._h, _ptr(be), _ptr(sb), L.LOC_HOST, self._stream()))
        return dict(lliks=ll, alpha=al, mx=mx, cs=cs, beta=be, sb=sb, logZ=lz[:, 0], lb_q4=lz[:, 1])

    @staticmethod
    def log_tables(loc, b=0):
        """The reference's unnormalised log-domain tables of window b from the scaled ones:
        lalpha[t] = log alpha[t] + sum_{u<=t}(log cs[u] + mx[u])          (hmmsgd_metaobs.py:775-803)
        lbeta[t]  = log beta[t]  + sum_{u=t}^{T-2}(log sb[u] + mx[u+1])   (:828-855)
        (float64 sums of the float32 scale factors; an entry whose normalised message is below the
        float32 range comes back as -inf)."""
        with np.errstate(divide='ignore'):
            cum = np.cumsum(np.log(loc["cs"][b].astype(np.float64)) + loc["mx"][b])
            lalpha = np.log(loc["alpha"][b].astype(np.float64)) + cum[:, None]
            d = np.log(loc["sb"][b].astype(np.float64))
            d[:-1] += loc["mx"][b][1:]
            d[-1] = 0.
            lbeta = np.log(loc["beta"][b].astype(np.float64)) + np.cumsum(d[::-1])[::-1][:, None]
        return lalpha, lbeta

    def ffbs(self, var_init, T=None, start=0, nsamples=1, seed=0):
        """hmm_fast.FFBS (hmm_fast.pyx:43-124): nsamples state paths (nsamples, T) int32 for the window
        obs[start:start+T] (default: the whole series)."""
        T = self.T_full if T is None else int(T)
        vi = _f64(var_init)
        z = np.empty((int(nsamples), T), dtype=np.int32)
        L.check(self.lib.svihmm_ffbs(self._h, _ptr(vi), int(start), T, int(nsamples), int(seed), _ptr(z),
                                     L.LOC_HOST, self._stream()))
        return z

    def svi_run(self, starts_all, T, tau, kappa, it0, bfact_A, bfact_E, flags=0, var_x=None, stats=None,
                peers=False):
        """nsteps global steps (E-step + natural-gradient update, hmmsgd_metaobs.py:396-439) enqueued by
        one C call.  starts_all: (nsteps, B) int64 CUDA tensor of window starts, drawn up front.
        Returns the stats tensor of the last step (device)."""
        assert isinstance(starts_all, torch.Tensor) and starts_all.is_cuda and starts_all.dtype == torch.int64
        starts_all = starts_all.contiguous()
        nsteps, B = int(starts_all.shape[0]), int(starts_all.shape[1])
        if stats is None:
            stats = self.new_stats()
        L.check(self.lib.svihmm_svi_run(self._h, _ptr(starts_all), nsteps, B, int(T), _ptr(var_x), _ptr(stats),
                                        int(flags), float(tau), float(kappa), int(it0), float(bfact_A),
                                        float(bfact_E), int(bool(peers)), self._stream()))
        self._keep["starts_all"] = starts_all
        return stats

    def global_bound(self, include_init=False):
        """Global part of the lower bound from the device-resident parameters
        (hmmsgd_metaobs.py:273-296; include_init: + the initial-distribution Dirichlet terms, hmmbase.py:145-199)."""
        out = np.empty(1)
        L.check(self.lib.svihmm_global_bound(self._h, _ptr(out), int(bool(include_init)), L.LOC_HOST, self._stream()))
        return float(out[0])

    def set_tuning(self, key, value):
        """svihmm_set_tuning (e.g. L.TUNE_B16_MIN_B: minibatch size from which the batched tensor-core
        path for K <= 16 is used; 0 = never)."""
        L.check(self.lib.svihmm_set_tuning(self._h, int(key), int(value)))

    def launch_count(self):
        return int(self.lib.svihmm_launch_count(self._h))

    def set_profiling(self, on=True):
        L.check(self.lib.svihmm_set_profiling(self._h, int(bool(on))))

    def phase_ms(self):
        """{phase name: (summed device ms, intervals)} since the last call (waits for the events)."""
        ms = (C.c_double * L.N_PHASES)()
        cnt = (C.c_int64 * L.N_PHASES)()
        L.check(self.lib.svihmm_get_phase_ms(self._h, ms, cnt))
        return {self.lib.svihmm_phase_name(i).decode(): (ms[i], int(cnt[i])) for i in range(L.N_PHASES)
                if cnt[i]}

    # ------------------------------------------------------------------ packing helpers
    def unpack_stats(self, stats):
        s = stats.detach().cpu().numpy() if isinstance(stats, torch.Tensor) else np.asarray(stats)
        K, D, DD, KE = self.K, self.D, self.DD, self.KE
        o = 0
        out = {}
        out["A"] = s[o:o + K * K].reshape(K, K); o += K * K
        out["n"] = s[o:o + KE]; o += KE
        out["sx"] = s[o:o + KE * D].reshape(KE, D); o += KE * D
        out["sxx"] = s[o:o + KE * DD].reshape({"niw_full": (KE, D, D), "niw_diag": (KE, D), "categorical": (KE, 0)}[self.emission]); o += KE * DD
        out["q0"] = s[o:o + K]; o += K
        out["logZ"], out["lb_q4"], out["B"] = float(s[o]), float(s[o + 1]), int(round(s[o + 2]))
        return out

    def pack_emit(self, mu, sigma, kappa, nu):
        """(K,D), (K,D,D)|(K,D), (K,)|(K,D), (K,)|(K,D) -> (K, plen) float64."""
        K, D = self.KE, self.D
        if self.emission == "categorical":              # mu = alpha_mf (K, C); the rest is ignored
            return _f64(mu).reshape(K, D).copy()
        out = np.empty((K, self.plen))
        mu = _f64(mu).reshape(K, D)
        if self.emission == "niw_full":
            out[:, :D] = mu
            out[:, D:D + D * D] = _f64(sigma).reshape(K, D * D)
            out[:, D + D * D] = _f64(kappa).reshape(K)
            out[:, D + D * D + 1] = _f64(nu).reshape(K)
        else:
            out[:, :D] = mu
            out[:, D:2 * D] = _f64(sigma).reshape(K, D)
            out[:, 2 * D:3 * D] = np.broadcast_to(_f64(kappa).reshape(K, -1), (K, D))
            out[:, 3 * D:] = np.broadcast_to(_f64(nu).reshape(K, -1), (K, D))
        return out

    def unpack_emit(self, em):
        K, D = self.KE, self.D
        em = np.asarray(em).reshape(K, self.plen)
        if self.emission == "categorical":
            return dict(alpha=em.copy())
        if self.emission == "niw_full":
            return dict(mu=em[:, :D].copy(), sigma=em[:, D:D + D * D].reshape(K, D, D).copy(),
                        kappa=em[:, D + D * D].copy(), nu=em[:, D + D * D + 1].copy())
        return dict(mu=em[:, :D].copy(), sigma=em[:, D:2 * D].copy(), kappa=em[:, 2 * D:3 * D].copy(),
                    nu=em[:, 3 * D:].copy())
