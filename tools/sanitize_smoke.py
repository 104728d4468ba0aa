"""Small pass over every kernel family for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py [n_points]
three tree configurations (order 7 in 3-D takes the streaming TMA M2L kernel, the others the grouped one) through every
evaluate entry point, the symmetric-P2P matvec, and one fit with a global trend."""
import numpy as np, sys
sys.path.insert(0, ".")
import ferreus_rbf_rs_b200 as fb
rng = np.random.default_rng(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
for dim, kt, order, nrhs in [(3, fb.FmmKernelType.LinearRbf, 7, 1), (2, fb.FmmKernelType.ThinPlateSplineRbf, 6, 4), (3, fb.FmmKernelType.CubicRbf, 9, 2)]:
    n = N
    centres = rng.random((6, dim))
    pts = np.ascontiguousarray(centres[rng.integers(0, 6, n)] + 0.03 * rng.standard_normal((n, dim)))
    w = rng.random((n, nrhs))
    t = fb.FmmTree(pts, order, fb.KernelParams(kt), True, True, params=fb.FmmParams(24, fb.M2LCompressionType.ACA, 1e-5, 1024))
    t.set_weights(w)
    a = t.evaluate(w, pts)
    b = t.evaluate(w, np.ascontiguousarray(pts[::-1]))
    idx = np.sort(rng.choice(n, n // 3, replace=False)).astype(np.uint64)
    c = t.evaluate_at_sources(w, idx)
    v, g = t.evaluate_with_gradients(w, pts[:500])
    print(dim, order, nrhs, float(np.abs(np.asarray(a).reshape(n, -1) - np.asarray(b).reshape(n, -1)[::-1]).max()), float(np.abs(np.asarray(a).reshape(n,-1)[idx.astype(int)] - np.asarray(c).reshape(len(idx),-1)).max()))
ic = fb.interpolant_config
pts = rng.random((min(N, 5000), 3)); vals = np.sin(3 * pts[:, 0]) + pts[:, 1]
m = fb.RBFInterpolator(pts, vals, ic.InterpolantSettings(ic.RBFKernelType.Linear), params=fb.config.Params(ic.RBFKernelType.Linear, ddm_params=fb.config.DDMParams(256, 0.5, 0.125, 600), naive_solve_threshold=100), global_trend=fb.GlobalTrend.three(20, 30, 10, 2, 1.5, 1))
print("fit ok", m.info()["iterations"], float(np.abs(m.evaluate_at_source() - vals).max()))
