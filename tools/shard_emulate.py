"""Per-rank kernel times of an N-GPU partition measured on ONE GPU (fb_tree_shard_as, csrc/comm.cu): for every rank of
`world` the share it would own is run with a world-1 communicator (collectives degenerate to copies), so the table shows
what each rank's kernels cost without paying for N GPUs.  Headline workload (BASELINE H) unless --n / --clustered.

    python tools/shard_emulate.py [--world 8] [--n 1000000] [--steps 5]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, nargs="+", default=[8])
    ap.add_argument("--n", type=int, default=1_000_000)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    import ferreus_rbf_rs_b200 as fb
    rng = np.random.default_rng(1000)
    pts = rng.random((args.n, 3))
    w = rng.random((args.n, 1))
    tree = fb.FmmTree(pts, 7, fb.KernelParams(fb.FmmKernelType.LinearRbf), True, True)
    tree.set_timing(True)
    tree.upload_weights(w)
    for _ in range(3):
        tree.matvec_resident()
    full = tree.last_timing()
    comm = fb.Communicator(0, 1)
    out = {"n": args.n, "unpartitioned_ms": full, "worlds": {}}
    for world in args.world:
        rows = []
        for r in range(world):
            tree.shard_as(comm, r, world)
            acc = []
            for it in range(3 + args.steps):
                tree.matvec_sharded()
                if it >= 3:
                    t = tree.sharded_timing()
                    t.update({"k_" + k: v for k, v in tree.last_timing().items()})
                    acc.append(t)
            med = {k: float(np.median([a[k] for a in acc])) for k in acc[0]}
            a, b = tree.shard_rows(r)
            med["rows"] = b - a
            med["step_ms"] = med["upward_exchange"] + med["downward"] + med["near_field_join_l2p"] + med["result_allreduce"]
            rows.append(med)
        out["worlds"][str(world)] = rows
        tree.shard(None)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
