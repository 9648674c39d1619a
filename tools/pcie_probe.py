"""Concurrent host<->device copy bandwidth of all ranks of one box (torchrun, one rank per GPU): what the platform gives the
zero-copy host path of bench.py's e2e figure when every GPU moves its queries and records over PCIe at the same time.
Prints one JSON line on rank 0: per-rank and aggregate GB/s for H2D, D2H and both directions at once."""
import json
import os

import torch
import torch.distributed as dist


def main():
    rank, world, local = (int(os.environ.get(k, "0")) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    world = max(world, 1)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = 64 << 20
    h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in, d_out = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
    s2 = torch.cuda.Stream(dev)

    def run(kind, reps=20):
        def once():
            if kind in ("h2d", "both"):
                d_in.copy_(h_in, non_blocking=True)
            if kind == "d2h":
                h_out.copy_(d_out, non_blocking=True)
            if kind == "both":
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        for _ in range(3):
            once()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            once()
        s = torch.cuda.current_stream(dev)
        s.wait_stream(s2)
        b.record()
        torch.cuda.synchronize(dev)
        ms = a.elapsed_time(b)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        nbytes = n * reps * (2 if kind == "both" else 1)
        return nbytes / (float(t.item()) * 1e-3) / 1e9

    out = {"n_gpus": world, "bytes_per_copy": n}
    for kind in ("h2d", "d2h", "both"):
        per = run(kind)
        out[kind] = {"per_gpu_GBps": per, "aggregate_GBps": per * world}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
