"""Where a multi-GPU step's time goes (torchrun, one rank per GPU): the scoring kernel alone, the kernel with the fused
all-gather stores but no barrier, the barrier alone, and the full step — CUDA events, max over ranks."""
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    rank, world, local = (int(os.environ.get(k, "0")) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from diffco_b200 import DiffCo, _lib
    from diffco_b200 import distributed as D
    from diffco_b200 import functional as Fn
    from diffco_b200 import kernel as K
    from diffco_b200 import model as M

    lib = _lib.load()
    B = bench.BATCH
    S, w, q = bench.make_problem(rank, B)
    robot = M.RevolutePlanarRobot(1.0, 0.3, dof=bench.DOF)
    chk = DiffCo(kernel_func=K.RQKernel(bench.GAMMA), transform=robot.fkine)
    chk.support_points = S.float().to(dev)
    chk.support_transformed = robot.fkine(chk.support_points)
    chk.gains = w.float().to(dev)
    sc = D.ShardedScorer(chk, weights="gains", group=dist.group.WORLD)
    qd = q.float().to(dev)
    sc.score_and_grad(qd)  # sets up the peer exchange
    peer = sc._peer
    sv, kfun = chk._select("gains")
    fk = chk._fk_for(sv)
    stream = Fn._stream_ptr(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def bcast_only():
        k = peer.steps & 1
        _lib.check(lib.dc_score_grad_bcast(C.byref(fk), C.byref(kfun.desc), C.byref(sv.desc), qd.data_ptr(), B, C.byref(peer.outs[k]),
                                           world, rank * B, _lib.DC_GRAD_SUM, None, stream), "bcast")
        peer.steps += 1

    def timed(fn, steps=50, pre=None):
        for _ in range(5):
            fn()
        torch.cuda.synchronize(dev)
        dist.barrier()
        tot = 0.0
        for _ in range(steps):
            flush.zero_()
            sc.align()
            if pre is not None:
                pre()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize(dev)
            tot += a.elapsed_time(b)
        t = torch.tensor([tot / steps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) * 1e3

    out = {"n_gpus": world, "unit": "us per step, max over ranks",
           "kernel_local": timed(lambda: sc.local_score_and_grad(qd)),
           "kernel_with_peer_stores": timed(bcast_only),
           "barrier_alone": timed(peer.barrier),
           "barrier_after_kernel": timed(peer.barrier, pre=bcast_only),
           "full_step": timed(lambda: sc.score_and_grad(qd))}
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
