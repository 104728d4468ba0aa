"""BASELINE.json matvec configs at full size: build, time (CUDA events around the resident matvec, host wall around the
reference-facing calls) and check against exact dense summation on sampled targets + linearity in the weights.

    python tools/config_bench.py C1|C2|C5|H [--n POINTS] [--steps K]

C1: 3-D LinearRbf, N = 100k uniform, order 6, 1 RHS.        C2: 2-D ThinPlateSpline, N = 1M uniform, order 9, 4 RHS.
C5: 3-D Spheroidal3, N = 10M clustered (512 blobs), order 7, eps 1e-7, 8 RHS.   H: 3-D LinearRbf 1M uniform, order 7.
"""
import argparse
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import ferreus_rbf_rs_b200 as fb  # noqa: E402
from oracle import kernels as okern  # noqa: E402


def clustered(rng, n, dim, n_centres):
    centres = rng.random((n_centres, dim))
    return centres[rng.integers(0, n_centres, n)] + 0.02 * rng.standard_normal((n, dim))


CONFIGS = {
    # name: (n, dim, generator, kernel type, oracle kernel index, order, eps, nrhs, seed)
    "C1": (100_000, 3, "uniform", "LinearRbf", 0, 6, 1e-6, 1, 42),
    "C2": (1_000_000, 2, "uniform", "ThinPlateSplineRbf", 1, 9, 1e-9, 4, 0),
    "C5": (10_000_000, 3, "clustered512", "SpheroidalRbf", 3, 7, 1e-7, 8, 0),
    "H": (1_000_000, 3, "uniform", "LinearRbf", 0, 7, 1e-7, 1, 1000),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config")
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--sample", type=int, default=200)
    args = ap.parse_args()
    n, dim, gen, kname, kidx, order, eps, nrhs, seed = CONFIGS[args.config]
    n = args.n or n
    rng = np.random.default_rng(seed)
    pts = rng.random((n, dim)) if gen == "uniform" else clustered(rng, n, dim, 512)
    pts = np.ascontiguousarray(pts)
    w = rng.random((n, nrhs))
    if kname == "SpheroidalRbf":
        kp = fb.KernelParams(fb.FmmKernelType.SpheroidalRbf, spheroidal_order=fb.SpheroidalOrder.Three,
                             base_range=1.0, total_sill=1.0)
    else:
        kp = fb.KernelParams(fb.FmmKernelType[kname])
    params = fb.FmmParams(256, fb.M2LCompressionType.ACA, eps, 1024)
    t0 = time.perf_counter()
    tree = fb.FmmTree(pts, order, kp, True, True, params=params)
    build_s = time.perf_counter() - t0
    info = tree.info()
    tree.set_timing(True)
    tree.upload_weights(w)
    for _ in range(2):
        tree.matvec_resident()
    dev_ms, stages = [], []
    for _ in range(args.steps):
        tree.matvec_resident()
        dev_ms.append(tree.last_matvec_ms())
        stages.append(tree.last_timing())
    t0 = time.perf_counter()
    tree.set_weights(w)
    y = np.asarray(tree.evaluate(w, pts)).reshape(n, nrhs)
    e2e_s = time.perf_counter() - t0
    # accuracy against exact summation on sampled targets; linearity in the weights
    sample = rng.integers(0, n, args.sample)
    dense = okern.dense_matvec(okern.Kernel(kidx, 1.0, 1.0), pts[sample], pts, w, block=64)
    err_dense = float(np.linalg.norm(y[sample] - dense) / np.linalg.norm(dense))
    w2 = rng.random((n, nrhs)) - 0.5
    tree.set_weights(w2)
    y2 = np.asarray(tree.evaluate(w2, pts)).reshape(n, nrhs)
    w3 = 2.0 * w - 3.0 * w2
    tree.set_weights(w3)
    y3 = np.asarray(tree.evaluate(w3, pts)).reshape(n, nrhs)
    lin = float(np.linalg.norm(y3 - (2.0 * y - 3.0 * y2)) / np.linalg.norm(y3))
    ms = float(np.median(dev_ms))
    med = {k: float(np.median([s[k] for s in stages])) for k in stages[0]}
    print(json.dumps({"config": args.config, "n": n, "dim": dim, "kernel": kname, "order": order, "nrhs": nrhs,
                      "tree": {"build_s": build_s, "depth": info["depth"], "cells": info["n_cells"],
                               "leaves": info["n_leaves"], "p2p_pairs": info["p2p_pairs"],
                               "m2p_pairs": info["m2p_pairs"], "v_entries": info["n_v"]},
                      "matvec_ms_resident": ms, "mpts_per_s_resident": n / ms / 1e3,
                      "e2e_s_set_weights_plus_evaluate": e2e_s, "mpts_per_s_e2e": n / e2e_s / 1e6,
                      "stages_ms": med, "rel_l2_vs_dense_sample": err_dense, "linearity_rel_l2": lin}))


if __name__ == "__main__":
    main()
