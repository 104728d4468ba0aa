"""Full RBF fit wall time (BASELINE.json config C3 recipe): clustered 3-D points, linear kernel, tol 1e-6,
default Params.  Usage: python tools/fit_bench.py N [kernel] [reps]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import ferreus_rbf_rs_b200 as fb  # noqa: E402


def f1_3d(p):  # smooth analytic test function (stand-in for rbf_test_functions.rs:102 f1_3d)
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    return 0.75 * np.exp(-((9 * x - 2) ** 2 + (9 * y - 2) ** 2 + (9 * z - 2) ** 2) / 4) + \
        0.5 * np.exp(-((9 * x - 7) ** 2 + (9 * y - 3) ** 2 + (9 * z - 5) ** 2) / 4)


def main():
    n = int(sys.argv[1])
    kernel = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    rng = np.random.default_rng(0)
    centres = rng.random((64, 3))
    pts = centres[rng.integers(0, 64, n)] + 0.02 * rng.standard_normal((n, 3))
    vals = f1_3d(pts)
    ic = fb.interpolant_config
    # warm the CUDA context and the lazily loaded kernels so the timed fit measures the algorithm
    wp = rng.random((30000, 3))
    fb.RBFInterpolator(wp, f1_3d(wp), ic.InterpolantSettings(ic.RBFKernelType(kernel)))
    events = []
    walls = []
    for rep in range(reps):
        events.clear()
        t0 = time.perf_counter()
        model = fb.RBFInterpolator(pts, vals, ic.InterpolantSettings(ic.RBFKernelType(kernel)),
                                   progress_callback=fb.progress.Progress(lambda e: events.append(e)))
        walls.append(time.perf_counter() - t0)
    wall = min(walls)
    info = model.info()
    res = [e.residual for e in events if isinstance(e, fb.progress.SolverIteration)]
    t1 = time.perf_counter()
    at_src = model.evaluate_at_source()
    t_eval = time.perf_counter() - t1
    err = float(np.linalg.norm(at_src - vals) / np.linalg.norm(vals))
    print(json.dumps({"n": n, "kernel": kernel, "fit_wall_s": wall, "fit_wall_s_all": walls, "setup_s": info["setup_seconds"],
                      "solve_s": info["solve_seconds"], "iterations": info["iterations"], "matvecs": info["matvecs"],
                      "ddm_domains": info["ddm_domains"], "last_residual": info["last_residual"],
                      "residual_history": res, "evaluate_at_source_s": t_eval, "fit_rel_l2_at_sources": err}))


if __name__ == "__main__":
    main()
