"""Shared fixtures: seeded point clouds and oracle/product construction with identical parameters."""
import numpy as np

from oracle import bbfmm as obb
from oracle import chebyshev as ocheb
from oracle import kernels as okern


def make_points(n, dim, kind, seed):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        pts = rng.random((n, dim))
    else:  # clustered: Gaussian blobs (BASELINE.md C3 recipe, fewer centres)
        centres = rng.random((8, dim))
        pts = centres[rng.integers(0, 8, n)] + 0.02 * rng.standard_normal((n, dim))
    return np.ascontiguousarray(pts)


def oracle_tree(pts, order, kernel_index, adaptive, sparse, max_pts, compression, eps, extents=None,
                base_range=1.0, total_sill=1.0):
    k = okern.Kernel(kernel_index, base_range, total_sill)
    return obb.FmmTree(pts, order, k, adaptive, sparse, extents,
                       obb.FmmParams(max_pts, compression, eps, 1024))


# registry index -> (FmmKernelType name, spheroidal order name)
PRODUCT_KERNEL = {0: ("LinearRbf", None), 1: ("ThinPlateSplineRbf", None), 2: ("CubicRbf", None),
                  3: ("SpheroidalRbf", "Three"), 4: ("SpheroidalRbf", "Five"), 5: ("SpheroidalRbf", "Seven"),
                  6: ("SpheroidalRbf", "Nine"), 7: ("Laplacian", None), 8: ("OneOverR2", None),
                  9: ("OneOverR4", None)}


def product_tree(pts, order, kernel_index, adaptive, sparse, max_pts, compression, eps, extents=None,
                 base_range=1.0, total_sill=1.0):
    import ferreus_rbf_rs_b200 as fb
    name, sph = PRODUCT_KERNEL[kernel_index]
    kp = fb.KernelParams(fb.FmmKernelType[name],
                         spheroidal_order=fb.SpheroidalOrder[sph] if sph else None,
                         base_range=base_range, total_sill=total_sill)
    params = fb.FmmParams(max_pts, fb.M2LCompressionType(compression), eps, 1024)
    return fb.FmmTree(pts, order, kp, adaptive, sparse, extents=extents, params=params)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def oracle_cell_table(ot):
    """(sorted keys, leaf flag per key) of an oracle tree."""
    keys = np.array(sorted(ot.lists.tree), dtype=np.uint64)
    flags = np.array([1 if int(k) in ot.lists.leaves else 0 for k in keys], dtype=np.uint8)
    return keys, flags


def product_lists_as_key_sets(pt, which, keys):
    ptr, idx = pt.dump_list(which)
    out = {}
    for c in range(len(keys)):
        a, b = int(ptr[c]), int(ptr[c + 1])
        if b > a:
            out[int(keys[c])] = set(int(keys[i]) for i in idx[a:b])
    return out
