import sys, json
import numpy as np
sys.path.insert(0, ".")
import ferreus_rbf_rs_b200 as fb
n = 1_000_000
rng = np.random.default_rng(0)
centres = rng.random((64, 3))
pts = centres[rng.integers(0, 64, n)] + 0.02 * rng.standard_normal((n, 3))
w = rng.random((n, 1))
tree = fb.FmmTree(pts, 7, fb.KernelParams(fb.FmmKernelType.LinearRbf), True, True)
tree.upload_weights(w)
tree.set_timing(True)
for m in (1984, 15872, 126976, n):
    if m < n:
        idx = np.sort(rng.choice(n, m, replace=False)).astype(np.uint64)
        tree.set_target_subset(idx)
    else:
        tree.set_target_subset(None)
    for _ in range(3):
        tree.matvec_resident()
    print(m, json.dumps({k: round(v, 3) for k, v in tree.last_timing().items()}))
