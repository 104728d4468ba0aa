"""Pin the oracle (and, on a GPU box, the CUDA path) against the REAL reference through the dump harness.

    python oracle/ref_harness/compare.py make   DIR     # writes the fixed inputs (raw little-endian f64)
    # ... on a machine with cargo, for every case printed by `make`:
    #     cargo run --release --manifest-path oracle/ref_harness/Cargo.toml -- matvec <args as printed>
    python oracle/ref_harness/compare.py check  DIR     # compares ref_*.f64 with the oracle (and the product if a GPU is present)

Test infrastructure only; never imported by the package.  No reference output is committed: the reference cannot be
built in this image (no Rust toolchain), which is why DESIGN.md calls parity "unpinned by reference vectors"."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

# name: (n, dim, kernel index, order, adaptive, sparse, max_pts, compression, eps, nrhs, seed)
CASES = {
    "c1_small": (20_000, 3, 0, 6, 1, 1, 256, 2, 1e-6, 1, 42),
    "tps_2d": (20_000, 2, 1, 9, 1, 1, 256, 2, 1e-9, 4, 0),
    "cubic_svd": (8_000, 3, 2, 5, 1, 1, 64, 1, 1e-5, 2, 7),
    "sph3_dense": (8_000, 3, 3, 5, 1, 0, 64, 0, 1e-5, 1, 8),
    "uniform_tree": (8_000, 3, 0, 5, 0, 1, 64, 2, 1e-5, 1, 9),
}


def inputs(case):
    n, dim, *_rest, nrhs, seed = CASES[case]
    rng = np.random.default_rng(seed)
    return rng.random((n, dim)), rng.random((n, nrhs))


def main():
    mode, d = sys.argv[1], sys.argv[2]
    os.makedirs(d, exist_ok=True)
    if mode == "make":
        for case, (n, dim, k, order, ad, sp, mx, comp, eps, nrhs, _seed) in CASES.items():
            pts, w = inputs(case)
            pts.astype("<f8").tofile(os.path.join(d, f"{case}_pts.f64"))
            w.astype("<f8").tofile(os.path.join(d, f"{case}_w.f64"))
            print(f"matvec {d}/{case}_pts.f64 {n} {dim} {d}/{case}_w.f64 {nrhs} {k} {order} {ad} {sp} {mx} {comp} {eps} "
                  f"{d}/ref_{case}.f64")
        return
    from oracle import bbfmm as obb, kernels as okern
    have_gpu = False
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        pass
    for case, (n, dim, k, order, ad, sp, mx, comp, eps, nrhs, _seed) in CASES.items():
        path = os.path.join(d, f"ref_{case}.f64")
        if not os.path.exists(path):
            print(case, "no reference dump")
            continue
        ref = np.fromfile(path, dtype="<f8").reshape(n, nrhs)
        pts, w = inputs(case)
        ot = obb.FmmTree(pts, order, okern.Kernel(k, 1.0, 1.0), bool(ad), bool(sp), None, obb.FmmParams(mx, comp, eps, 1024))
        ot.set_weights(w)
        got = ot.evaluate(w, pts)
        rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        line = f"{case}: oracle vs reference rel-L2 {rel:.3e} (gate 1e-10)"
        if have_gpu:
            from tests import helpers as H
            pt = H.product_tree(pts, order, k, bool(ad), bool(sp), mx, comp, eps)
            pt.set_weights(w)
            g = np.asarray(pt.evaluate(w, pts)).reshape(n, nrhs)
            line += f"; CUDA vs reference {np.linalg.norm(g - ref) / np.linalg.norm(ref):.3e}"
        print(line)


if __name__ == "__main__":
    main()
