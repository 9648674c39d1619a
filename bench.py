#!/usr/bin/env python
"""bench.py — collision score+grad evals/sec on the configuration BASELINE.json's metric is quoted on.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

Workload (config.workload): BASELINE.json configs[1] — 7-DoF planar arm (RevolutePlanarRobot.fkine), FK + RQKernel(10),
2000 support vectors, batch 65536 per GPU, score + gradient w.r.t. the configuration, float32, synthetic inputs
(seed 1234).  One step = one pass of the hot path over one batch.  N > 1 (torchrun, one rank per GPU): every rank
scores its own 65536-row shard of the global batch (weak scaling) straight into its slice of the gathered
[score | grad] buffer and one NCCL all-gather makes the whole batch visible on every rank, inside the timed region.

Timing: W untimed warm-up steps, then K steps, each bracketed by CUDA events on the launching stream with the L2
flushed (a 256 MiB memset) before every step; `value` = evals of all ranks / max-over-ranks summed step time.
`e2e` runs the public host-buffer API (pinned host q -> H2D -> fused kernel -> D2H of [score | grad]) per step.
`--impl reference` times the reference algorithm's CPU port (oracle/) on the host cores instead.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "collision score+grad evals/sec (7-DoF, 2k SVs)"
UNIT = "evals/s"
N_SV, DOF, N_FEAT, N_CLASS, BATCH = 2000, 7, 14, 1, 65536
GAMMA = 10.0
SEED = 1234
BYTES_PER_EVAL = 4 * (2 * DOF + 2 * N_CLASS)            # read q, write score, read grad_out (implicit ones), write grad
FLOPS_PER_EVAL = N_SV * (5 * N_FEAT + 8 + 4 * N_CLASS)
WORKLOAD = "configs[1]: 7-DoF planar arm, FK+RQKernel(gamma=10,p=2), 2000 SVs, batch 65536/GPU, score+grad, fp32"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="configurations per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tc", type=int, default=1, choices=[0, 1],
                    help="1: tensor-core kernel (default); 0: force the FP32-pipe thread-per-query kernel")
    return ap.parse_args()


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(p.get("sm_max_mhz", 1965.0))
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def make_problem(rank, batch):
    """Synthetic cfg-2 inputs, CPU float64 from explicit seeds."""
    import math

    import torch

    gen = torch.Generator().manual_seed(SEED)
    S = (torch.rand(N_SV, DOF, generator=gen, dtype=torch.float64) * 2 - 1) * math.pi
    w = torch.randn(N_SV, generator=gen, dtype=torch.float64)
    gq = torch.Generator().manual_seed(SEED + 1 + rank)
    q = (torch.rand(batch, DOF, generator=gq, dtype=torch.float64) * 2 - 1) * math.pi
    return S, w, q


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference algorithm's CPU port (oracle/) on the host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_rate(S, w, q, budget_s, chunk=8192, min_chunks=1):
    """evals/s of the oracle (reference algorithm: FK -> cdist -> RQ -> matmul, autograd backward) in float32 on all
    host threads, over as many `chunk`-row slices of the workload as fit in `budget_s` seconds."""
    import torch

    from oracle import diffco_oracle as O

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    L = torch.ones(DOF, dtype=torch.float32)
    fk = lambda z: O.fk_planar_chain(z, L)
    St = fk(S.float())
    wf = w.float()
    kern = O.KernelSpec("rq", GAMMA, 2)
    f = lambda z: O.score_original(z, fk, kern, St, wf)
    qf = q.float()
    O.score_and_grad(f, qf[:chunk])  # warm-up
    done, t0 = 0, time.perf_counter()
    i = 0
    while True:
        lo = (i * chunk) % max(1, len(qf) - chunk + 1)
        O.score_and_grad(f, qf[lo:lo + chunk])
        done += min(chunk, len(qf) - lo)
        i += 1
        el = time.perf_counter() - t0
        if i >= min_chunks and el >= budget_s:
            break
    return done / el, threads, done, el


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    S, w, q = make_problem(0, args.batch)
    chunk = 8192
    # each step = one bounded sample of the workload (8192 configurations); the whole run is capped to ~3 minutes
    import torch

    from oracle import diffco_oracle as O

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    L = torch.ones(DOF, dtype=torch.float32)
    fk = lambda z: O.fk_planar_chain(z, L)
    St, wf, qf = fk(S.float()), w.float(), q.float()
    kern = O.KernelSpec("rq", GAMMA, 2)
    f = lambda z: O.score_original(z, fk, kern, St, wf)
    t0 = time.perf_counter()
    O.score_and_grad(f, qf[:chunk])
    probe = time.perf_counter() - t0
    while chunk > 512 and probe * (args.steps + args.warmup) * (chunk / 8192) > 150.0:
        chunk //= 2
    for i in range(args.warmup):
        O.score_and_grad(f, qf[(i * chunk) % (len(qf) - chunk + 1):][:chunk])
    t0 = time.perf_counter()
    for i in range(args.steps):
        O.score_and_grad(f, qf[(i * chunk) % (len(qf) - chunk + 1):][:chunk])
    el = time.perf_counter() - t0
    value = args.steps * chunk / el
    sample = f"{chunk} of the {args.batch} configurations per step, float32, torch CPU {threads} threads"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": WORKLOAD, "n_sv": N_SV, "dof": DOF, "batch_per_gpu": args.batch, "sample_per_step": chunk},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}
    BAD = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        get_reasons = getattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(
            self.nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = get_reasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def visible_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ----------------------------------------------------------------------------------------------------------------
# native arm
# ----------------------------------------------------------------------------------------------------------------
def run_native(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl native needs a CUDA device: diffco_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    from diffco_b200 import DiffCo, _lib
    from diffco_b200 import distributed as D
    from diffco_b200 import kernel as K
    from diffco_b200 import model as M

    lib = _lib.load()
    _lib.check(lib.dc_set_option(_lib.DC_OPT_TC_ENABLE, float(args.tc)), "dc_set_option")
    B = args.batch
    S, w, q = make_problem(rank, B)
    robot = M.RevolutePlanarRobot(1.0, 0.3, dof=DOF)
    checker = DiffCo(kernel_func=K.RQKernel(GAMMA), transform=robot.fkine)
    checker.support_points = S.float().to(dev)
    checker.support_transformed = robot.fkine(checker.support_points)  # dc_fk_forward on the device
    checker.gains = w.float().to(dev)
    scorer = D.ShardedScorer(checker, weights="gains", group=(dist.group.WORLD if world > 1 else None))

    q_dev = q.float().to(dev)
    q_host = q.float().pin_memory()
    out_host = torch.empty((B, N_CLASS + DOF), dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(step_fn, steps, warmup):
        for _ in range(warmup):
            step_fn()
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        sampler = ClockSampler(visible_index(local))
        sampler.start()
        l0 = lib.dc_launch_count()
        t0 = time.perf_counter()
        for a, b in evs:
            flush.zero_()  # evict the previous step's inputs/outputs from the 126 MB L2 (untimed)
            if world > 1:
                scorer.align()  # untimed: the flushes end at different moments on different GPUs; without this the wait
                                # for the slowest flush would be charged to the step through the step's own collective
            a.record(stream)
            step_fn()
            b.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        launches = lib.dc_launch_count() - l0
        clocks = sampler.stop()
        ms = sum(a.elapsed_time(b) for a, b in evs)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, clocks, wall

    # ---- device-resident throughput (`value`) -------------------------------------------------------------
    def step_dev():
        return scorer.score_and_grad(q_dev)

    def measure(fn):
        res = timed(fn, args.steps, args.warmup)
        if any(r in ClockSampler.BAD for r in res[2]["reasons"]):  # rejected: re-measure once
            res = timed(fn, args.steps, max(3, args.warmup))
        return res

    ms_dev, launches, clocks, _ = measure(step_dev)
    value = world * B * args.steps / (ms_dev * 1e-3)

    # ---- the gathered result is right: the block another rank contributed equals a local evaluation of ITS shard ------
    gather_check = None
    if world > 1:
        peer = (rank + 1) % world
        _, _, q_peer = make_problem(peer, B)
        s_all, g_all = scorer.score_and_grad(q_dev)
        got = torch.cat([s_all, g_all], dim=1)[peer * B:(peer + 1) * B].clone()
        s_loc, g_loc = scorer.local_score_and_grad(q_peer.float().to(dev))
        want = torch.cat([s_loc, g_loc], dim=1)
        bad = torch.tensor([0 if torch.equal(got, want) else 1], device=dev)
        dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        gather_check = "ok" if int(bad.item()) == 0 else "MISMATCH"
        if gather_check != "ok":
            raise SystemExit("bench.py: gathered records differ from a local evaluation of the peer's shard")

    # ---- dominant kernel alone (roofline): the fused score+grad launch without the collective ----------------
    def step_kernel():
        return scorer.local_score_and_grad(q_dev)

    ms_k, k_launches, _, _ = timed(step_kernel, args.steps, 3)
    per_launch_s = ms_k * 1e-3 / max(1, k_launches)
    which = lib.dc_last_score_kernel()
    kernel_name = {2: "score_tc_kernel<GRAD> (tcgen05 kind::f16, TMEM, bulk TMA; RQ2, C=1, F=14)",
                   1: "score_tq_kernel<F=14,RQ2,C=1,GRAD> (FP32 pipe, bulk TMA)"}.get(which, _lib.KERNEL_NAMES.get(which, "?"))
    hbm_peak, peak_src, sm_max = load_peaks()
    achieved_gbs = BYTES_PER_EVAL * B / per_launch_s / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            traffic = json.load(f).get("tensor-core" if which == 2 else "thread-per-query", {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    fp32_peak_tflops = 148 * 128 * 2 * sm_max * 1e6 / 1e12
    achieved_tflops = FLOPS_PER_EVAL * B / per_launch_s / 1e12
    # the unit that bounds the tensor-core kernel: one MUFU reciprocal per (query, support) pair, 16 per clock per SM
    pairs_per_clk_sm = B * N_SV / per_launch_s / (148 * sm_max * 1e6)

    # ---- end to end through the host-buffer API (`e2e`) -----------------------------------------------------
    def step_e2e():
        return scorer.score_and_grad_host(q_host, out_host)

    ms_e2e, _, _, _ = timed(step_e2e, args.steps, max(3, args.warmup // 2))
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)

    # ---- reference CPU path beside it (rank 0, N == 1 only; bounded sample) -----------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, threads, done, el = cpu_reference_rate(S, w, q, budget_s=12.0)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{done} configurations ({el:.1f} s) of the same workload in 8192-row chunks, float32, torch CPU"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_sv": N_SV, "dof": DOF, "n_features": N_FEAT, "batch_per_gpu": B,
                       "global_batch": world * B, "kernel": "RQKernel(gamma=10,p=2)",
                       "sharding": ("none" if world == 1 else
                                    f"batch rows over {world} ranks; all-gather of [score|grad] " +
                                    ("fused into the kernel epilogue (peer stores over NVLink + flag barrier)"
                                     if getattr(scorer, "_peer", None) else "by one NCCL all_gather_into_tensor")),
                       "gather_check": gather_check,
                       "arithmetic": ("fp32 in / out; radial profile, score and near pairs in fp32; the two contractions as tcgen05 "
                                      "kind::f16 MMAs on operands split into two 11-bit terms (products ~22 bits) with fp32 "
                                      "accumulation; parity vs the float64 oracle equals the FP32-pipe kernel's (1.7e-6 score, "
                                      "1.8e-6 grad of max on this workload, gate 1e-5; tests/test_gpu_tc.py)"
                                      if which == 2 else "fp32 throughout (packed FP32 pipe)"),
                       "l2": "flushed before every step (256 MiB memset, untimed" +
                             ("; ranks re-aligned by an untimed device barrier after it" if world > 1 else "") +
                             "); per-step CUDA events summed"},
            "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved_gbs / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": kernel_name, "kernel_ms": per_launch_s * 1e3,
                         "algorithmic_bytes_per_launch": BYTES_PER_EVAL * B,
                         "note": "the path is compute bound (2600 flop/B), not HBM bound: see roofline_compute"},
            "roofline_compute": ({"bound": "mufu (one reciprocal per query-support pair; contractions on tcgen05)",
                                  "achieved": pairs_per_clk_sm, "peak": 16.0, "unit": "pairs/clk/SM",
                                  "frac": pairs_per_clk_sm / 16.0,
                                  "peak_source": f"16 MUFU lanes per SM at {sm_max:.0f} MHz (nominal max clock)",
                                  "algorithmic_tflops": achieved_tflops,
                                  "fp32_fma_pipe_peak_tflops": fp32_peak_tflops} if which == 2 else
                                 {"bound": "fp32_fma_pipe", "achieved": achieved_tflops, "peak": fp32_peak_tflops,
                                  "unit": "TFLOP/s", "frac": achieved_tflops / fp32_peak_tflops,
                                  "peak_source": f"148 SMs x 128 lanes x 2 flop x {sm_max:.0f} MHz (nominal max clock)",
                                  "algorithmic_flops_per_launch": FLOPS_PER_EVAL * B}),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": q_host.numel() * 4,
                    "d2h_bytes_per_step": out_host.numel() * 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
