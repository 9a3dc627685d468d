"""CPU-side checks of the drop-in boundary: libsvihmm.so loads, exports every symbol that
include/svihmm.h declares (and the ctypes table binds exactly those), and fails loudly - not
silently on a CPU path - when there is no CUDA device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "svihmm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(svihmm_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    syms = header_symbols()
    for must in ("svihmm_create", "svihmm_destroy", "svihmm_set_series", "svihmm_set_globals",
                 "svihmm_estep", "svihmm_estep_host", "svihmm_global_update", "svihmm_batch_update",
                 "svihmm_get_locals", "svihmm_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from pysvihmm_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build with python -m pysvihmm_b200.build"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in header_symbols():
        assert hasattr(lib, s), "libsvihmm.so does not export %s" % s


def test_ctypes_table_matches_header():
    from pysvihmm_b200 import _lib
    assert sorted(_lib.SYMBOLS) == header_symbols()
    lib = _lib.load()
    assert lib.svihmm_version() >= 100


def test_no_cpu_fallback():
    """Without a CUDA device creating a context must fail with a message, and the Python engine
    must refuse to construct; nothing falls back to the oracle or to numpy."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("box has a GPU")
    from pysvihmm_b200 import SvihmmError, _lib
    from pysvihmm_b200.engine import EStepEngine
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.svihmm_create(ctypes.byref(h), 0, 4, 2, _lib.EMIT_NIW_FULL)
    assert rc == _lib.ECUDA and b"cuda" in lib.svihmm_last_error().lower()
    with pytest.raises(SvihmmError):
        EStepEngine(4, 2)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pysvihmm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
