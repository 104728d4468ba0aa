"""What the two exchanges of the partitioned matvec cost on this box (run under torchrun, one rank per GPU):
all-reduce of the multipole array (4M doubles at the 1M-point headline) and of the result (1M doubles), alone and under
a compute kernel that fills the SMs, against an all-to-all of the same multipole volume by grouped send / recv."""
import json
import os
import sys

import torch
import torch.distributed as dist


def timed(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def main():
    rank = int(os.environ["RANK"])
    lr = int(os.environ["LOCAL_RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    mult = torch.ones(4_004_504, dtype=torch.float64, device="cuda")
    res = torch.ones(1_000_000, dtype=torch.float64, device="cuda")
    out = {}
    out["allreduce_mult_32MB_alone_ms"] = timed(lambda: dist.all_reduce(mult))
    out["allreduce_result_8MB_alone_ms"] = timed(lambda: dist.all_reduce(res))
    small = torch.ones(40 * 344, dtype=torch.float64, device="cuda")
    out["allreduce_110KB_alone_ms"] = timed(lambda: dist.all_reduce(small))
    # all-to-all of the same volume: every rank sends its eighth to every peer
    share = mult.numel() // world
    recv = torch.empty(world * share, dtype=torch.float64, device="cuda")
    out["all_gather_mult_32MB_alone_ms"] = timed(lambda: dist.all_gather_into_tensor(recv, mult[rank * share:(rank + 1) * share]))

    def sendrecv():
        ops = []
        for p in range(world):
            if p == rank:
                continue
            ops.append(dist.P2POp(dist.isend, mult[rank * share:(rank + 1) * share], p))
            ops.append(dist.P2POp(dist.irecv, recv[p * share:(p + 1) * share], p))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    out["sendrecv_all_to_all_32MB_alone_ms"] = timed(sendrecv)
    # the same collectives under a kernel that fills every SM (FP64 GEMM on a second stream)
    a = torch.randn(4096, 4096, dtype=torch.float64, device="cuda")
    side = torch.cuda.Stream()

    def under(fn):
        def run():
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                torch.mm(a, a)
            fn()
            torch.cuda.current_stream().wait_stream(side)
        return run
    out["gemm_alone_ms"] = timed(under(lambda: None))
    out["allreduce_mult_32MB_under_gemm_ms"] = timed(under(lambda: dist.all_reduce(mult)))
    out["all_gather_mult_under_gemm_ms"] = timed(under(lambda: dist.all_gather_into_tensor(recv, mult[rank * share:(rank + 1) * share])))
    out["sendrecv_under_gemm_ms"] = timed(under(sendrecv))
    if rank == 0:
        print(json.dumps(out), file=sys.stderr)
        json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "s2_nccl_probe.json"), "w"))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
