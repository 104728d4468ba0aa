"""Partitioned BBFMM matvec across the GPUs of one node (csrc/comm.cu): correctness against the unpartitioned matvec on the
same tree, per-rank per-stage device times, whole-step time (max over ranks).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/sharded_bench.py H|C5 [--n POINTS] [--steps K] [--no-check]

H: 3-D LinearRbf, 1M uniform, order 7, 1 RHS (BASELINE headline).  C5: 3-D Spheroidal3, clustered (512 blobs), order 7,
eps 1e-7, 8 RHS (BASELINE config 5; --n sets the cloud size, default 10M).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("config", choices=["H", "C5"])
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--fork-mode", type=int, default=-1, help="near-field fork point (fb_tree_shard_fork_mode); -1 = library default")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    out_fd = os.dup(1)
    os.dup2(2, 1)  # NCCL banners go to stderr
    dist.init_process_group("nccl" if world > 1 else "gloo", rank=rank, world_size=world,
                            **({"device_id": torch.device("cuda", local_rank)} if world > 1 else {}))
    import ferreus_rbf_rs_b200 as fb
    from ferreus_rbf_rs_b200 import _lib
    _lib.lib().fb_set_device(local_rank)

    if args.config == "H":
        n = args.n or 1_000_000
        rng = np.random.default_rng(1000)
        pts = rng.random((n, 3))
        w = rng.random((n, 1))
        kp, order, eps = fb.KernelParams(fb.FmmKernelType.LinearRbf), 7, 1e-7
    else:
        n = args.n or 10_000_000
        rng = np.random.default_rng(0)
        centres = rng.random((512, 3))
        pts = np.ascontiguousarray(centres[rng.integers(0, 512, n)] + 0.02 * rng.standard_normal((n, 3)))
        w = rng.random((n, 8))
        kp = fb.KernelParams(fb.FmmKernelType.SpheroidalRbf, spheroidal_order=fb.SpheroidalOrder.Three)
        order, eps = 7, 1e-7
    t0 = time.perf_counter()
    tree = fb.FmmTree(pts, order, kp, True, True, params=fb.FmmParams(256, fb.M2LCompressionType.ACA, eps, 1024))
    build_s = time.perf_counter() - t0
    comm = fb.Communicator.from_torch_distributed(dist)
    tree.upload_weights(w)  # before the cut: the work model depends on the number of right-hand sides
    tree.shard(comm)
    if args.fork_mode >= 0:
        tree.shard_fork_mode(args.fork_mode)
    rows = [tree.shard_rows(r) for r in range(world)]

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        tree.matvec_sharded()
    barrier()
    walls, stages = [], []
    for _ in range(args.steps):
        barrier()
        t0 = time.perf_counter()
        tree.matvec_sharded()
        walls.append(time.perf_counter() - t0)
        stages.append(tree.sharded_timing())
    barrier()
    med = {k: float(np.median([s[k] for s in stages])) for k in stages[0]}
    dev_ms = float(np.median([sum(s.values()) for s in stages]))
    err = None
    if not args.no_check:
        got = np.array(tree.sharded_download())
        tree.matvec_resident()
        ref = np.array(tree.download_result())
        err = float(np.linalg.norm(got - ref) / np.linalg.norm(ref))
    mine = torch.tensor([dev_ms, 1e3 * float(np.median(walls)), med["upward_exchange"], med["downward"],
                         med["near_field_join_l2p"], med["result_allreduce"], float(rows[rank][1] - rows[rank][0]),
                         err if err is not None else -1.0], dtype=torch.float64,
                        device="cuda" if world > 1 else "cpu")
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    if rank == 0:
        tab = np.stack([v.cpu().numpy() for v in allv])
        step_ms = float(tab[:, 0].max())
        os.dup2(out_fd, 1)
        print(json.dumps({
            "config": args.config, "fork_mode": args.fork_mode, "n": n, "nrhs": int(w.shape[1]), "n_gpus": world, "tree_build_s": build_s,
            "ms_per_matvec_device_max_over_ranks": step_ms, "ms_per_matvec_wall_max_over_ranks": float(tab[:, 1].max()),
            "mpts_per_s": n / (step_ms * 1e-3) / 1e6,
            "per_rank": [{"rank": r, "rows": int(tab[r, 6]), "device_ms": tab[r, 0], "upward_exchange": tab[r, 2],
                          "downward": tab[r, 3], "near_field_join_l2p": tab[r, 4], "result_allreduce": tab[r, 5]}
                         for r in range(world)],
            "rel_l2_vs_unpartitioned": [float(v) for v in tab[:, 7]] if err is not None else None}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
