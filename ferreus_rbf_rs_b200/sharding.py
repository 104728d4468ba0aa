"""Host-side sharding of one BBFMM tree by Morton-contiguous leaf ranges: the evaluate-a-subset variant
driven from Python, kept for the CPU (gloo) tests of the cut and for callers without NCCL.  The production multi-GPU path
is the library's own partition (csrc/comm.cu: FmmTree.shard / matvec_sharded, fb_comm_*), where no kernel evaluation is
made twice and the exchanges are NCCL all-reduces on the device.

Every rank holds the same tree (points and interaction lists are replicated; the upward pass, 2 % of a
matvec, is recomputed locally).  The leaves, in Morton order, are cut into `world` contiguous ranges balanced
by estimated work (direct pairs + M2L entries, not point counts — clustered clouds are very uneven), and a
rank evaluates only the targets of its own range.  The one exchange step is the all-gather of the result
slices (NCCL on GPUs, gloo in the CPU tests); there is no collective inside a pass.
"""
import numpy as np


def partition_by_work(work, n_parts):
    """Cut `work` (per leaf, Morton order) into n_parts contiguous ranges with nearly equal sums.
    Returns n_parts + 1 non-decreasing boundaries into the leaf sequence."""
    work = np.asarray(work, dtype=np.float64)
    n = work.size
    if n_parts <= 1 or n == 0:
        return np.array([0, n], dtype=np.int64) if n_parts <= 1 else np.linspace(0, n, n_parts + 1).astype(np.int64)
    csum = np.cumsum(work)
    total = csum[-1]
    targets = total * np.arange(1, n_parts) / n_parts
    cuts = np.searchsorted(csum, targets, side="left") + 1
    bounds = np.concatenate([[0], np.minimum(cuts, n), [n]]).astype(np.int64)
    return np.maximum.accumulate(bounds)


def rank_rows(leaf_ptr, order, bounds, rank):
    """source rows (targets) owned by `rank`: the points of its contiguous leaf range, in Morton order"""
    a, b = int(leaf_ptr[bounds[rank]]), int(leaf_ptr[bounds[rank + 1]])
    return np.ascontiguousarray(order[a:b])


class ShardedMatvec:
    """y = A w with the rows of y split across ranks.  `tree` needs leaf_work(), morton_order() and either
    evaluate_at_sources(w, idx) (host arrays) or the resident trio upload_weights / matvec_resident /
    download_result with set_target_subset."""

    def __init__(self, tree, rank, world):
        self.tree, self.rank, self.world = tree, rank, world
        leaf_ptr, work = tree.leaf_work()
        self.order = tree.morton_order()
        self.bounds = partition_by_work(work, world)
        self.rows = [rank_rows(leaf_ptr, self.order, self.bounds, r) for r in range(world)]
        self.my_rows = self.rows[rank]
        self.n = int(leaf_ptr[-1])

    def local(self, w):
        """this rank's slice of A w (rows self.my_rows)"""
        if self.my_rows.size == 0:
            return np.zeros((0,) + tuple(np.shape(w)[1:]))
        return np.asarray(self.tree.evaluate_at_sources(w, self.my_rows))

    def gather(self, local, dist=None):
        """all-gather the slices into the full result on every rank (torch.distributed, any backend)"""
        if dist is None or self.world == 1:
            out = np.zeros((self.n,) + local.shape[1:])
            out[self.my_rows] = local
            return out
        import torch
        # NCCL moves device memory only: stage the slices on this rank's GPU there, on the host for gloo / mpi
        dev = (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl"
               else torch.device("cpu"))
        sizes = [r.size for r in self.rows]
        width = int(np.prod(local.shape[1:])) if local.ndim > 1 else 1
        pad = max(sizes) * width
        buf = torch.zeros(pad, dtype=torch.float64, device=dev)
        buf[: local.size] = torch.from_numpy(np.ascontiguousarray(local).reshape(-1)).to(dev)
        bufs = [torch.zeros(pad, dtype=torch.float64, device=dev) for _ in range(self.world)]
        dist.all_gather(bufs, buf)
        out = np.zeros((self.n, width))
        for r in range(self.world):
            out[self.rows[r]] = bufs[r][: sizes[r] * width].cpu().numpy().reshape(sizes[r], width)
        return out if local.ndim > 1 else out[:, 0]
