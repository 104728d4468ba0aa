"""Host-only tree + interaction lists (include/ferreus_b200.h, fb_host_tree_*) against the oracle, bit for bit, without a
GPU: Morton keys of every cell, leaf set, leaf membership (ascending source rows) and the U / V / W / X lists as sets of
keys (morton.rs:29-373, linear_tree.rs:20-485).  The device path builds the same HostTree from codes sorted by the device
radix sort (tests/test_gpu_fmm.py::test_tree_and_lists_bit_exact checks that one on the GPU)."""
import ctypes as C

import numpy as np
import pytest

from ferreus_rbf_rs_b200 import _lib
from oracle import linear_tree, morton
from tests import helpers as H

U64P = C.POINTER(C.c_uint64)


def host_tree(pts, max_pts, adaptive, sparse, extents=None):
    L = _lib.lib()
    n, dim = pts.shape
    h = C.c_void_p()
    ext = None if extents is None else np.ascontiguousarray(extents, dtype=np.float64)
    rc = L.fb_host_tree_new(_lib.dptr(pts), n, dim, dim, 1, None if ext is None else _lib.dptr(ext), max_pts,
                            1 if adaptive else 0, 1 if sparse else 0, C.byref(h))
    assert rc == 0
    nc, nl, depth = C.c_uint64(), C.c_uint64(), C.c_int32()
    nlist = (C.c_uint64 * 4)()
    assert L.fb_host_tree_counts(h, C.byref(nc), C.byref(nl), C.byref(depth), nlist) == 0
    nc = int(nc.value)
    keys = np.zeros(nc, dtype=np.uint64)
    flags = np.zeros(nc, dtype=np.uint8)
    ptr = np.zeros(nc + 1, dtype=np.uint64)
    idx = np.zeros(n, dtype=np.uint64)
    assert L.fb_host_tree_dump_cells(h, keys.ctypes.data_as(U64P), flags.ctypes.data_as(C.POINTER(C.c_uint8)),
                                     ptr.ctypes.data_as(U64P), idx.ctypes.data_as(U64P)) == 0
    lists = []
    for which in range(4):
        lp = np.zeros(nc + 1, dtype=np.uint64)
        li = np.zeros(max(int(nlist[which]), 1), dtype=np.uint64)
        assert L.fb_host_tree_dump_list(h, which, lp.ctypes.data_as(U64P), li.ctypes.data_as(U64P)) == 0
        lists.append((lp, li[:int(nlist[which])]))
    L.fb_host_tree_free(h)
    return int(depth.value), keys, flags, ptr, idx, lists


CASES = [
    # n, dim, kind, adaptive, sparse, max_pts
    (6000, 3, "uniform", True, True, 40),
    (5000, 3, "clustered", True, True, 25),
    (3000, 3, "clustered", True, False, 25),
    (5000, 2, "clustered", True, True, 20),
    (2000, 1, "uniform", True, True, 16),
    (4000, 3, "clustered", False, True, 30),
    (4000, 2, "uniform", False, False, 30),
    (300, 3, "uniform", True, True, 256),      # root split only
]


@pytest.mark.parametrize("n,dim,kind,adaptive,sparse,max_pts", CASES)
def test_host_tree_and_lists_bit_exact(n, dim, kind, adaptive, sparse, max_pts):
    pts = H.make_points(n, dim, kind, seed=11)
    if kind == "clustered":  # exact duplicates and a ragged tail
        pts[-30:] = pts[:30]
    lo, hi = pts.min(axis=0), pts.max(axis=0)
    center, radius = morton.calculate_tree_center_and_radius(list(lo) + list(hi))
    ot = linear_tree.build_tree(pts, center, radius, max_pts, not sparse, dim, adaptive)
    depth, keys, flags, ptr, idx, lists = host_tree(pts, max_pts, adaptive, sparse)
    assert depth == ot.depth
    order = np.argsort(keys)
    okeys = np.array(sorted(ot.tree), dtype=np.uint64)
    oflags = np.array([1 if int(k) in ot.leaves else 0 for k in okeys], dtype=np.uint8)
    assert np.array_equal(keys[order], okeys), "cell key sets differ"
    assert np.array_equal(flags[order], oflags), "leaf sets differ"
    for c in range(len(keys)):
        if flags[c]:
            mine = idx[int(ptr[c]):int(ptr[c + 1])]
            ref = ot.leaf_source_indices.get(int(keys[c]), np.zeros(0, dtype=np.int64))
            assert np.array_equal(mine.astype(np.int64), np.asarray(ref, dtype=np.int64))
    for which, ref in enumerate([ot.u_lists, ot.v_lists, ot.w_lists or {}, ot.x_lists or {}]):
        lp, li = lists[which]
        mine = {}
        for c in range(len(keys)):
            a, b = int(lp[c]), int(lp[c + 1])
            if b > a:
                mine[int(keys[c])] = set(int(keys[i]) for i in li[a:b])
        ref = {k: set(v) for k, v in ref.items() if len(v) > 0}
        assert mine == ref, f"list {'UVWX'[which]} differs"


def test_host_tree_extents_and_outside_points():
    """user extents move the root cube (morton.rs:349-373); a source outside them is refused, as fb_tree_new does"""
    pts = H.make_points(2000, 3, "uniform", seed=5)
    ext = [-1.0, -1.0, -1.0, 2.0, 2.0, 2.0]
    center, radius = morton.calculate_tree_center_and_radius(ext)
    ot = linear_tree.build_tree(pts, center, radius, 30, False, 3, True)
    depth, keys, flags, _, _, _ = host_tree(pts, 30, True, True, extents=ext)
    assert depth == ot.depth
    assert np.array_equal(np.sort(keys), np.array(sorted(ot.tree), dtype=np.uint64))
    L = _lib.lib()
    h = C.c_void_p()
    small = np.array([0.0, 0.0, 0.0, 0.5, 0.5, 0.5])
    far = np.ascontiguousarray(pts + 10.0)
    rc = L.fb_host_tree_new(_lib.dptr(far), len(far), 3, 3, 1, _lib.dptr(small), 30, 1, 1, C.byref(h))
    assert rc == _lib.FB_ERR_INVALID_ARGUMENT
