"""Where the end-to-end matvec time goes (host buffers in, host buffers out): per-call wall times of the public API at the
bench workload, next to the resident matvec, and the same C-ABI calls with a reused (already faulted-in) output buffer."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ferreus_rbf_rs_b200 as fb  # noqa: E402
from ferreus_rbf_rs_b200 import _lib  # noqa: E402


def med(f, n=7):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        f()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


def main():
    n = int(os.environ.get("N", 1_000_000))
    rng = np.random.default_rng(1)
    pts = rng.random((n, 3))
    ws = [rng.random((n, 1)) for _ in range(2)]
    tree = fb.FmmTree(pts, 7, fb.KernelParams(fb.FmmKernelType.LinearRbf), True, True)
    k = [0]

    def nxt():
        k[0] ^= 1
        return ws[k[0]]

    for _ in range(3):
        w = nxt(); tree.set_weights(w); tree.evaluate(w, pts)
    res = {}
    res["set_weights_ms"] = med(lambda: tree.set_weights(nxt()))
    w = nxt(); tree.set_weights(w)
    res["evaluate_same_w_ms"] = med(lambda: tree.evaluate(w, pts))

    def full():
        w = nxt(); tree.set_weights(w); tree.evaluate(w, pts)
    res["set_weights_plus_evaluate_ms"] = med(full)

    def at_src():
        w = nxt(); tree.set_weights(w); tree.evaluate_at_sources(w)
    res["set_weights_plus_evaluate_at_sources_ms"] = med(at_src)
    tree.upload_weights(w)
    res["matvec_resident_ms"] = med(lambda: tree.matvec_resident())
    res["np_zeros_and_touch_8MB_ms"] = med(lambda: np.zeros((n, 1)).fill(1.0))
    # raw C ABI with a reused output buffer
    L = _lib.lib()
    out = np.zeros((n, 1))
    bad = C.c_uint64(0)

    def raw():
        w = nxt()
        L.fb_tree_set_weights(tree._h, _lib.dptr(w), n, 1, 1, 1)
        L.fb_tree_evaluate(tree._h, _lib.dptr(w), n, 1, 1, 1, _lib.dptr(pts), n, 3, 1, _lib.dptr(out), None, 1, 1,
                           C.byref(bad))
    raw(); raw()
    res["c_abi_reused_output_ms"] = med(raw)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
