"""Host-side sharding logic on CPU: work-balanced Morton-contiguous partition + all-gather of result slices,
world_size 2 over gloo.  The per-rank compute is stood in for by the oracle's dense summation."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from ferreus_rbf_rs_b200.sharding import ShardedMatvec, partition_by_work


def test_partition_by_work_balances_and_covers():
    rng = np.random.default_rng(0)
    work = rng.random(1000) ** 4 * 100 + 1
    for parts in (1, 2, 3, 8):
        b = partition_by_work(work, parts)
        assert b[0] == 0 and b[-1] == 1000 and np.all(np.diff(b) >= 0) and len(b) == parts + 1
        sums = [work[b[i]:b[i + 1]].sum() for i in range(parts)]
        assert max(sums) <= work.sum() / parts + work.max() + 1e-9
    b = partition_by_work(np.ones(4), 8)          # more ranks than leaves: some ranks own nothing
    assert b[0] == 0 and b[-1] == 4 and len(b) == 9 and np.all(np.diff(b) >= 0)


class _FakeTree:
    """stands in for FmmTree on a CPU-only machine: same sharding interface, dense oracle arithmetic"""

    def __init__(self, pts, leaf_size=37):
        from oracle import kernels as ok
        self.pts = pts
        self.kernel = ok.Kernel(ok.LINEAR)
        self.order_ = np.lexsort((pts[:, 2], pts[:, 1], pts[:, 0])).astype(np.int64)
        n = pts.shape[0]
        self.ptr = np.append(np.arange(0, n, leaf_size), n).astype(np.int64)

    def leaf_work(self):
        return self.ptr, np.diff(self.ptr).astype(np.float64) ** 2

    def morton_order(self):
        return self.order_

    def evaluate_at_sources(self, w, idx):
        from oracle import kernels as ok
        return ok.dense_matvec(self.kernel, self.pts[idx], self.pts, w)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    pts = rng.random((700, 3))
    w = rng.random((700, 2))
    sm = ShardedMatvec(_FakeTree(pts), rank, world)
    full = sm.gather(sm.local(w), dist)
    from oracle import kernels as ok
    ref = ok.dense_matvec(ok.Kernel(ok.LINEAR), pts, pts, w)
    ok_ = bool(np.allclose(full, ref, rtol=1e-13, atol=1e-13))
    covered = np.sort(np.concatenate(sm.rows))
    ok_ = ok_ and np.array_equal(covered, np.arange(700)) and 0 < sm.my_rows.size < 700
    q.put((rank, ok_))
    dist.destroy_process_group()


def test_sharded_matvec_world2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]
