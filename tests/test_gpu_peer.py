"""GPU test (-m gpu) of the fused all-gather: dc_score_grad_bcast stores every tile's [score | grad] records into every
rank's gathered buffer (CUDA IPC mappings) and dc_peer_barrier publishes the step.  Two processes share cuda:0 here (IPC
works within one device; the ranks talk over gloo), so the test runs on a single-GPU box; bench.py --gpus N is the
same code over NVLink."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret, rname="planar7", b=4224, fold="0"):
    import torch.distributed as dist

    from diffco_b200 import DiffCo, _lib
    from diffco_b200 import distributed as D
    from diffco_b200 import kernel as K
    from tests import problems as P

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["DIFFCO_B200_PEER_SYNC"] = fold  # "1": the step barrier in the tail of the scoring kernel (one launch)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        dev = torch.device("cuda", 0)
        robot, S, W = P.synthetic_model(rname, 700, 1, seed=77)
        dc = DiffCo(kernel_func=K.RQKernel(10.0), transform=robot.fkine)
        dc.support_points = S.float().to(dev)
        dc.support_transformed = robot.fkine(dc.support_points)
        dc.gains = W[:, 0].float().to(dev)
        scorer = D.ShardedScorer(dc, weights="gains", group=dist.group.WORLD)
        single = D.ShardedScorer(dc, weights="gains")
        worst = 0.0
        q_host = torch.empty((b, robot.dof), dtype=torch.float32).pin_memory()
        out_host = torch.empty((b, 1 + robot.dof), dtype=torch.float32).pin_memory()
        for step in range(5):  # both alternating buffers, several epochs
            q = P.sample_configs(robot, world * b, torch.Generator().manual_seed(100 + step)).float().to(dev)
            s, g = scorer.score_and_grad(q[rank * b:(rank + 1) * b])
            assert isinstance(scorer._peer, D.PeerExchange), "fused all-gather path was not taken"
            s, g = s.clone(), g.clone()
            s1, g1 = single.score_and_grad(q)
            assert _lib.load().dc_last_score_kernel() == 2
            worst = max(worst, float((s - s1).abs().max() / s1.abs().max()), float((g - g1).abs().max() / g1.abs().max()))
            # host buffers in / host buffers out with world > 1: ONE launch (pinned q read zero-copy, this rank's records
            # mirrored to the pinned output) — must equal this rank's rows of the single-process result
            q_host.copy_(q[rank * b:(rank + 1) * b])
            out_host.fill_(float("nan"))
            scorer.score_and_grad_host(q_host, out_host)
            torch.cuda.synchronize(dev)
            want = torch.cat([s1, g1], dim=1)[rank * b:(rank + 1) * b].cpu()
            assert torch.equal(out_host, want), float((out_host - want).abs().max())
        ret[rank] = worst
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("rname,b,fold", [("planar7", 4224, "0"), ("se2arm", 4099, "0"), ("planar7", 4224, "1"), ("se2arm", 4099, "1")])
def test_fused_all_gather_two_ranks_one_device(rname, b, fold, cuda_device):
    """planar7: 8-float records, 33 full tiles per rank (bulk-TMA peer stores).  se2arm: 7-float records and a ragged shard, so
    rank 1's block starts at an address that is NOT 16-byte aligned — the per-thread store path must take over.  fold = "1":
    dc_score_grad_bcast_sync (the barrier in the kernel's tail) instead of dc_score_grad_bcast + dc_peer_barrier."""
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret, rname, b, fold)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    for p in procs:
        if p.is_alive():
            p.terminate()
        assert p.exitcode == 0, f"rank exited with {p.exitcode}"
    # position independence of the kernel: the sharded result equals the single-process result bit for bit
    assert ret[0] == 0.0 and ret[1] == 0.0, dict(ret)
