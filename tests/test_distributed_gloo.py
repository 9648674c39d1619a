"""CPU tests of the multi-GPU host logic (diffco_b200/distributed.py) with the gloo backend, world_size 2 and 3:
row partitioning, padding of ragged batches, the in-place gathered [score | grad] buffer and trimming.  The local
scorer is a stub (the CUDA kernel needs a GPU; its numerics are the -m gpu tests' business)."""
import os
import socket
import types

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffco_b200 import distributed as D


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _stub_local(q, out):
    out[:, :2] = torch.stack([q.sum(1), (q**2).sum(1)], 1)
    out[:, 2:] = 3.0 * q


def _worker(rank, world, port, total, q_all, results):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        stub = types.SimpleNamespace(n_class=2, dof=q_all.shape[1], dtype=torch.float64, device=torch.device("cpu"))
        scorer = D.ShardedScorer(stub, group=dist.group.WORLD, local_fn=_stub_local)
        assert scorer.world == world and scorer.rank == rank
        # (1) replicated global batch, ragged: every rank ends up with all rows
        s, g = scorer.score_and_grad_global(q_all[:total])
        ok1 = (s.shape == (total, 2) and torch.allclose(s[:, 0], q_all[:total].sum(1)) and
               torch.allclose(s[:, 1], (q_all[:total] ** 2).sum(1)) and torch.allclose(g, 3.0 * q_all[:total]))
        # (2) per-rank shards of equal size b (bench.py's weak-scaling step)
        b = 5
        mine = q_all[rank * b:(rank + 1) * b].contiguous()
        s2, g2 = scorer.score_and_grad(mine)
        ref = q_all[:world * b]
        ok2 = s2.shape == (world * b, 2) and torch.allclose(s2[:, 0], ref.sum(1)) and torch.allclose(g2, 3.0 * ref)
        results[rank] = bool(ok1 and ok2)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 11), (2, 8), (3, 7), (2, 1), (3, 0)])
def test_sharded_scorer_gloo(world, total):
    q_all = torch.randn(32, 4, dtype=torch.float64, generator=torch.Generator().manual_seed(0))
    results = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(world, _free_port(), total, q_all, results), nprocs=world, join=True)
    assert all(results.get(r) for r in range(world)), dict(results)


def test_shard_bounds_cover_the_batch_exactly():
    for total in (0, 1, 7, 8, 65536, 2097152, 2097153):
        for world in (1, 2, 3, 4, 8):
            rows = []
            for r in range(world):
                lo, hi, per = D.shard_bounds(total, world, r)
                assert 0 <= hi - lo <= per
                rows += list(range(lo, hi)) if total < 100 else [(lo, hi)]
            if total < 100:
                assert rows == list(range(total))
            else:
                assert rows[0][0] == 0 and rows[-1][1] == total and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))


def _peer_failure_worker(rank, world, port, results):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        raised = False
        try:  # no CUDA device here: allocation fails on every rank -> they must all raise, in step, without hanging
            D.PeerExchange(64, 8, torch.float32, torch.device("cpu"), dist.group.WORLD)
        except RuntimeError:
            raised = True
        t = torch.tensor([rank + 1.0])
        dist.all_reduce(t)  # the ranks' collective sequences are still aligned
        results[rank] = bool(raised and t.item() == world * (world + 1) / 2)
    finally:
        dist.destroy_process_group()


def test_peer_exchange_setup_fails_consistently_without_peer_memory():
    """The fused all-gather's set-up (diffco_b200/distributed.py::PeerExchange) must degrade on ALL ranks together — the
    caller then uses NCCL everywhere — and keep the ranks' collectives in step whatever failed locally."""
    world = 2
    results = mp.get_context("spawn").Manager().dict()
    mp.spawn(_peer_failure_worker, args=(world, _free_port(), results), nprocs=world, join=True)
    assert all(results.get(r) for r in range(world)), dict(results)


def test_numa_binding_helper_leaves_the_process_alone_without_a_gpu():
    """bind_host_thread_to_gpu pins a rank to its GPU's NUMA-local CPUs before it allocates pinned buffers; when the
    topology cannot be read (no NVML / no GPU, as here) it must return None and change nothing."""
    import os

    from diffco_b200 import distributed as D

    before = os.sched_getaffinity(0)
    assert D.bind_host_thread_to_gpu(0) is None or os.sched_getaffinity(0) <= before
    if not torch.cuda.is_available():
        assert os.sched_getaffinity(0) == before
