"""ctypes binding of libsvihmm.so (include/svihmm.h).  Fails loudly when the CUDA library is
missing: there is no CPU fallback in this package."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SVIHMM_LIB", os.path.join(HERE, "lib", "libsvihmm.so"))   # override: A/B builds

OK, EINVAL, ECUDA, ENOMEM, ESTATE, EUNSUPPORTED = 0, -1, -2, -3, -4, -5
EMIT_NIW_FULL, EMIT_NIW_DIAG, EMIT_CATEGORICAL = 0, 1, 2
F32, F64 = 0, 1
LOC_DEVICE, LOC_HOST = 0, 1
WRAP, ADD_PRIOR, MASK_LL, EXACT_XI, KEEP_LOCALS, BF16_DENSE = 1, 2, 4, 8, 16, 32
N_PHASES = 8
TUNE_B16_MIN_B = 1
TUNE_SCAN_MIN_T = 2
TUNE_NO_HOSTREG = 3

_vp, _i, _i64, _d, _u = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_uint
# name -> (restype, argtypes); must list every symbol include/svihmm.h declares
SYMBOLS = {
    "svihmm_last_error": (C.c_char_p, []),
    "svihmm_version": (_i, []),
    "svihmm_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i]),
    "svihmm_create_mix": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i]),
    "svihmm_set_mix_weights": (_i, [_vp, _vp, _vp, _i, _vp]),
    "svihmm_get_mix_weights": (_i, [_vp, _vp, _i, _vp]),
    "svihmm_destroy": (_i, [_vp]),
    "svihmm_emit_param_len": (C.c_size_t, [_vp]),
    "svihmm_stats_len": (C.c_size_t, [_vp]),
    "svihmm_set_series": (_i, [_vp, _vp, _i64, _i, _vp, _i, _vp]),
    "svihmm_set_series_streamed": (_i, [_vp, _vp, _i64, _i, _vp]),
    "svihmm_set_prior": (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    "svihmm_set_globals": (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    "svihmm_get_globals": (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    "svihmm_estep": (_i, [_vp, _vp, _i, _i, _vp, _vp, _u, _vp]),
    "svihmm_estep_buffered": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _u, _vp]),
    "svihmm_set_var_init": (_i, [_vp, _vp, _i, _vp]),
    "svihmm_estep_host": (_i, [_vp, _vp, _i, _i, _vp, _vp, _u, _vp]),
    "svihmm_prefetch_windows": (_i, [_vp, _vp, _i, _i]),
    "svihmm_estep_streamed": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _u, _vp]),
    "svihmm_svi_step_host": (_i, [_vp, _vp, _i, _i, _vp, _vp, _u, _d, _d, _d, _vp]),
    "svihmm_global_update": (_i, [_vp, _vp, _d, _d, _d, _vp]),
    "svihmm_comm_buffer_len": (C.c_size_t, [_vp]),
    "svihmm_comm_attach": (_i, [_vp, _i, _i, _vp]),
    "svihmm_global_update_peers": (_i, [_vp, _vp, _d, _d, _d, _vp]),
    "svihmm_get_reduced_stats": (_i, [_vp, _vp, _i, _vp]),
    "svihmm_set_adagrad": (_i, [_vp, _i, _vp]),
    "svihmm_batch_update": (_i, [_vp, _vp, _vp]),
    "svihmm_batchsgd_update": (_i, [_vp, _vp, _d, _vp]),
    "svihmm_get_locals": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    "svihmm_svi_run": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _u, _d, _d, _i64, _d, _d, _i, _vp]),
    "svihmm_global_bound": (_i, [_vp, _vp, _i, _i, _vp]),
    "svihmm_check": (_i, [_vp, _vp]),
    "svihmm_set_tuning": (_i, [_vp, _i, _i]),
    "svihmm_get_locals_beta": (_i, [_vp, _vp, _vp, _i, _vp]),
    "svihmm_ffbs": (_i, [_vp, _vp, _i64, _i, _i, C.c_uint64, _vp, _i, _vp]),
    "svihmm_launch_count": (_i64, [_vp]),
    "svihmm_set_profiling": (_i, [_vp, _i]),
    "svihmm_get_phase_ms": (_i, [_vp, _vp, _vp]),
    "svihmm_phase_name": (C.c_char_p, [_i]),
}

_lib = None


class SvihmmError(RuntimeError):
    pass


def load():
    """dlopen the in-tree library and bind every exported symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SvihmmError(
            "%s not found: build it with `python -m pysvihmm_b200.build` (nvcc, sm_100a). "
            "pysvihmm_b200 has no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        msg = load().svihmm_last_error()
        raise SvihmmError("libsvihmm error %d: %s" % (rc, msg.decode() if msg else "?"))
