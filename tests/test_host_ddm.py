"""Host-only duplicate removal + DDM hierarchy (include/ferreus_rbf_b200.h, fr_host_ddm_*) against the oracle without a
GPU: kept rows (rbf.rs:1391-1467), level point sets, per-domain point sets, internal masks, and the ORDER of the points of
every domain (internal points, then overlap points by ascending point-to-box distance, domain_decomposition.rs:236-311;
Domain::factorise moves the special points to the front, domain.rs:219-260)."""
import ctypes as C

import numpy as np
import pytest

from ferreus_rbf_rs_b200 import _lib
from oracle import rbf as orbf
from tests import helpers as H

U64P = C.POINTER(C.c_uint64)


def host_ddm(pts, kernel, leaf, coarse, naive):
    L = _lib.lib()
    n, dim = pts.shape
    st = _lib.FrSettings()
    L.fr_settings_default(kernel, C.byref(st))
    pr = _lib.FrParams()
    L.fr_params_default(kernel, C.byref(pr))
    pr.leaf_threshold, pr.overlap_quota, pr.coarse_ratio, pr.coarse_threshold = leaf, 0.5, 0.125, coarse
    pr.naive_solve_threshold = naive
    h = C.c_void_p()
    assert L.fr_host_ddm_new(_lib.dptr(pts), n, dim, dim, 1, C.byref(st), C.byref(pr), C.byref(h)) == 0
    nk, nl = C.c_uint64(), C.c_uint64()
    assert L.fr_host_ddm_counts(h, C.byref(nk), C.byref(nl)) == 0
    kept = np.zeros(int(nk.value), dtype=np.uint64)
    assert L.fr_host_ddm_kept(h, kept.ctypes.data_as(U64P)) == 0
    levels = []
    for li in range(int(nl.value)):
        nd, npts = C.c_uint64(), C.c_uint64()
        assert L.fr_host_ddm_level(h, li, C.byref(nd), C.byref(npts), None, None, None, None) == 0
        lp = np.zeros(int(npts.value), dtype=np.uint64)
        ptr = np.zeros(int(nd.value) + 1, dtype=np.uint64)
        assert L.fr_host_ddm_level(h, li, None, None, lp.ctypes.data_as(U64P), ptr.ctypes.data_as(U64P), None, None) == 0
        idx = np.zeros(int(ptr[-1]), dtype=np.uint64)
        internal = np.zeros(int(ptr[-1]), dtype=np.uint8)
        assert L.fr_host_ddm_level(h, li, None, None, None, None, idx.ctypes.data_as(U64P),
                                   internal.ctypes.data_as(C.POINTER(C.c_uint8))) == 0
        levels.append((lp, ptr, idx, internal))
    L.fr_host_ddm_free(h)
    return kept.astype(np.int64), levels


@pytest.mark.parametrize("kernel,dim,n", [(0, 3, 2600), (2, 3, 2200), (1, 2, 2600), (3, 3, 1800)])
def test_host_ddm_hierarchy_sets_masks_and_order(kernel, dim, n):
    pts = H.make_points(n, dim, "clustered", seed=141)
    pts[-25:] = pts[:25]  # exact duplicates: removed before the hierarchy is built
    s = orbf.InterpolantSettings(kernel)
    s.set_basis_size(dim)
    keep = np.asarray(orbf.remove_duplicates(pts, s.kernel()), dtype=np.int64)
    kept, levels = host_ddm(pts, kernel, 128, 300, 100)
    assert np.array_equal(kept, keep) and len(keep) == n - 25
    ddm = orbf.DDMTree(pts[keep], s, 128, 0.5, 0.125, 300, factorise=False)
    assert len(levels) == len(ddm.levels)
    checked = 0
    for (lp, ptr, idx, internal), lvl in zip(levels, ddm.levels):
        assert np.array_equal(lp.astype(np.int64), lvl.point_indices)
        assert len(ptr) - 1 == len(lvl.leaf_domains)
        for d, dom in enumerate(lvl.leaf_domains):
            a, b = int(ptr[d]), int(ptr[d + 1])
            mine = idx[a:b].astype(np.int64).tolist()
            ref = [int(v) for v in dom.idx]
            assert sorted(mine) == sorted(ref)
            ok = False
            for rk in range(0, s.basis_size + 1):  # rk special points were moved to the front
                front = set(mine[:rk])
                if len(front) == rk and mine[rk:] == [v for v in ref if v not in front]:
                    ok = True
                    break
            assert ok, f"domain {d}: point order differs from the oracle's"
            k = min(len(dom.mask), len(dom.idx))
            ref_int = {int(v) for v, mk in zip(dom.idx[:k], dom.mask[:k]) if mk}
            assert {v for v, f in zip(mine, internal[a:b]) if f} == ref_int
            checked += 1
    assert checked >= 8


def test_host_ddm_small_problem_has_no_levels():
    pts = H.make_points(400, 3, "uniform", seed=3)
    kept, levels = host_ddm(pts, 0, 128, 300, 4096)
    assert len(kept) == 400 and levels == []
