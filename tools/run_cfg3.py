"""BASELINE configs[2] on its own (Baxter F = 12, 4 classes, 5000 SVs, Polyharmonic(1,1) rbf_score, 262144 queries): a few
launches of score_tq_kernel<F=12,PH1,C=4,GRAD>, for profiling under ncu (bench.py's configs.cfg3 times the same thing)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    from diffco_b200 import _lib

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    stream = torch.cuda.current_stream(dev)

    def timed(fn, steps, warmup, sc=None, do_flush=False):
        for _ in range(warmup):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(steps):
            fn()
        b.record(stream)
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b), steps, {}, 0.0

    print(bench.bench_cfg3(torch, lib, dev, timed, bench.load_peaks()[0]))


if __name__ == "__main__":
    main()
