"""The C-ABI library loads on a CPU-only machine and exports every symbol declared in include/*.h."""
import ctypes
import os
import re

from ferreus_rbf_rs_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for fn in os.listdir(inc):
        if fn.endswith(".h"):
            text = open(os.path.join(inc, fn)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            names |= set(re.findall(r"\b((?:fb|fr)_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in sorted(syms) if not hasattr(lib, s)]
    assert not missing, missing


def test_python_signatures_cover_the_header():
    assert declared_symbols() <= set(_lib.SIGNATURES) | set(getattr(_lib, "SOLVER_SIGNATURES", {}))


def test_invalid_arguments_are_reported_not_fatal():
    L = _lib.lib()
    out = ctypes.c_void_p()
    assert L.fb_tree_new(None, 0, 3, 3, 1, 5, None, 1, 1, None, None, ctypes.byref(out)) == _lib.FB_ERR_INVALID_ARGUMENT
    assert "source_points" in _lib.last_error()
    assert L.fb_tree_m2l_rank(None, 2, 0) == -1


def test_pinned_pool_falls_back_without_a_device():
    """_lib.pinned hands out ordinary numpy memory when fb_host_alloc cannot page-lock (no GPU in this container) and for
    small results; either way the array is writable and has the requested shape."""
    import numpy as np
    from ferreus_rbf_rs_b200 import _lib
    small = _lib.pinned.empty((10, 2))
    big = _lib.pinned.empty((200_000, 1))
    for a, shape in ((small, (10, 2)), (big, (200_000, 1))):
        assert a.shape == shape and a.dtype == np.float64 and a.flags.writeable and a.flags.c_contiguous
        a[:] = 1.0
