"""Small-batch latency of the persistent evaluator: build_evaluator once, then
evaluate_targets / evaluate_targets_with_gradients on batches of B targets — what ferreus_rmt's surface follower does
thousands of times (rbf.rs:1009-1042).   python tools/evaluator_latency.py [N] [B]"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import ferreus_rbf_rs_b200 as fb  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    rng = np.random.default_rng(0)
    pts = rng.random((n, 3))
    vals = np.sin(4 * pts[:, 0]) * np.cos(3 * pts[:, 1]) + pts[:, 2]
    ic = fb.interpolant_config
    t0 = time.perf_counter()
    model = fb.RBFInterpolator(pts, vals, ic.InterpolantSettings(ic.RBFKernelType.Linear))
    fit_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    model.build_evaluator([0.0, 0.0, 0.0, 1.0, 1.0, 1.0])
    build_s = time.perf_counter() - t0
    out = {"n": n, "batch": b, "fit_s": fit_s, "build_evaluator_s": build_s}
    for name, fn in (("evaluate_targets", model.evaluate_targets),):
        lat = []
        for it in range(60):
            tg = rng.random((b, 3))
            t0 = time.perf_counter()
            fn(tg)
            lat.append(time.perf_counter() - t0)
        lat = np.array(lat[10:])
        out[name] = {"median_ms": float(np.median(lat) * 1e3), "p90_ms": float(np.quantile(lat, 0.9) * 1e3),
                     "targets_per_s": float(b / np.median(lat))}
    tg = rng.random((b, 3))
    ref = model.evaluate(tg)
    got = model.evaluate_targets(tg)
    out["evaluator_vs_one_shot_rel_l2"] = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
