#!/usr/bin/env python
"""bench.py — collision score+grad evals/sec on the configuration BASELINE.json's metric is quoted on.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--extras 0|1]

Headline workload (config.workload): BASELINE.json configs[1] — 7-DoF planar arm (RevolutePlanarRobot.fkine), FK +
RQKernel(10), 2000 support vectors, batch 65536 per GPU, score + gradient w.r.t. the configuration, float32, synthetic
inputs (seed 1234).  One step = one pass of the hot path over one batch.  N > 1 (torchrun, one rank per GPU): every rank
scores its own 65536-row shard (weak scaling) and the [score | grad] records of all ranks are gathered on every rank
inside the timed region (fused into the kernel's epilogue over NVLink, or one NCCL all-gather).

Timing: W untimed warm-up steps, then K steps, each bracketed by CUDA events on the launching stream with the L2
flushed (a 256 MiB memset) before every step; `value` = evals of all ranks / max-over-ranks summed step time.
`e2e`: the public host-buffer API per step (pinned host q in, pinned host [score | grad] out).  `sustained`: the same
launch back to back for >= 2 s.  `configs`: the other BASELINE configurations this GPU path covers — cfg3 (Baxter, 4
classes, 5000 SVs, 262144 queries), cfg4 (256-waypoint trajectory step), cfg5 (2 097 152 queries split over the ranks:
the strong-scaling line).  `--impl reference` times the reference algorithm's CPU port (oracle/) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import math
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
REF_CHUNK = 8192  # rows per reference step: FIXED (the (B, N) temporaries of the reference algorithm bound it), so that
                  # the reference arm reads the same from run to run
N_SMS = 148


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="configurations per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extras", type=int, default=1, choices=[0, 1],
                    help="1 (default): also measure sustained / autograd / cfg3 / cfg4 / cfg5 / ATen-on-GPU; 0: headline only")
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


def load_profile_json(name):
    try:
        with open(os.path.join(ROOT, "profiles", name)) as f:
            return json.load(f)
    except Exception:
        return {}


def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def make_problem(rank, batch, n_sv=N_SV, dof=DOF):
    """Synthetic cfg-2 inputs, CPU float64 from explicit seeds."""
    import torch

    gen = torch.Generator().manual_seed(SEED)
    S = (torch.rand(n_sv, dof, generator=gen, dtype=torch.float64) * 2 - 1) * math.pi
    w = torch.randn(n_sv, generator=gen, dtype=torch.float64)
    gq = torch.Generator().manual_seed(SEED + 1 + rank)
    q = (torch.rand(batch, dof, generator=gq, dtype=torch.float64) * 2 - 1) * math.pi
    return S, w, q


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference algorithm's CPU port (oracle/) on the host cores
# ----------------------------------------------------------------------------------------------------------------
def _reference_fn(S, w, dtype, device="cpu"):
    import torch

    from oracle import diffco_oracle as O

    L = torch.ones(DOF, dtype=dtype, device=device)
    fk = lambda z: O.fk_planar_chain(z, L)
    St = fk(S.to(device=device, dtype=dtype))
    wf = w.to(device=device, dtype=dtype)
    kern = O.KernelSpec("rq", GAMMA, 2)
    f = lambda z: O.score_original(z, fk, kern, St, wf)
    return lambda z: O.score_and_grad(f, z)


def cpu_reference_rate(S, w, q, budget_s, dtype=None, chunk=REF_CHUNK):
    """evals/s of the oracle (reference algorithm: FK -> cdist -> RQ -> matmul, autograd backward) on all host threads,
    over as many `chunk`-row slices of the workload as fit in `budget_s` seconds."""
    import torch

    dtype = dtype or torch.float32
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    run = _reference_fn(S, w, dtype)
    qf = q.to(dtype)
    run(qf[:chunk])  # warm-up
    done, i, t0 = 0, 0, time.perf_counter()
    while True:
        lo = (i * chunk) % max(1, len(qf) - chunk + 1)
        run(qf[lo:lo + chunk])
        done += min(chunk, len(qf) - lo)
        i += 1
        el = time.perf_counter() - t0
        if el >= budget_s:
            break
    return done / el, threads, done, el


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch

    S, w, q = make_problem(0, args.batch)
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    run = _reference_fn(S, w, torch.float32)
    qf = q.float()
    span = max(1, len(qf) - REF_CHUNK + 1)
    for i in range(max(1, args.warmup)):
        run(qf[(i * REF_CHUNK) % span:][:REF_CHUNK])
    t0 = time.perf_counter()
    for i in range(args.steps):
        run(qf[(i * REF_CHUNK) % span:][:REF_CHUNK])
    el = time.perf_counter() - t0
    value = args.steps * REF_CHUNK / el
    rate64, _, done64, el64 = cpu_reference_rate(S, w, q, budget_s=4.0, dtype=torch.float64)
    sample = f"{REF_CHUNK} of the {args.batch} configurations per step (fixed), float32, torch CPU {threads} threads"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "impl": "reference",
        "config": {"workload": WORKLOAD, "n_sv": N_SV, "dof": DOF, "batch_per_gpu": args.batch, "sample_per_step": REF_CHUNK},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "cpu_model": cpu_model(),
                         "float64": {"value": rate64, "unit": UNIT,
                                     "sample": f"{done64} configurations ({el64:.1f} s), float64 (what the reference's scripts run)"}},
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
        self.index, self.samples, self.reasons, self.max_mhz, self.power = index, [], set(), None, []
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
                self.power.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
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
                "samples": len(s), "power_w_max": (max(self.power) if self.power else None)}


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

    from diffco_b200 import DiffCo, MultiDiffCo, _lib
    from diffco_b200 import distributed as D
    from diffco_b200 import kernel as K
    from diffco_b200 import model as M

    lib = _lib.load()
    _lib.check(lib.dc_set_option(_lib.DC_OPT_TC_ENABLE, float(args.tc)), "dc_set_option")
    # the pinned buffers of the host path are allocated below: put them on this GPU's memory node
    numa_cpus = D.bind_host_thread_to_gpu(visible_index(local)) if world > 1 else None
    group = dist.group.WORLD if world > 1 else None
    B = args.batch
    S, w, q = make_problem(rank, B)
    robot = M.RevolutePlanarRobot(1.0, 0.3, dof=DOF)
    checker = DiffCo(kernel_func=K.RQKernel(GAMMA), transform=robot.fkine)
    checker.support_points = S.float().to(dev)
    checker.support_transformed = robot.fkine(checker.support_points)  # dc_fk_forward on the device
    checker.gains = w.float().to(dev)
    scorer = D.ShardedScorer(checker, weights="gains", group=group)

    q_dev = q.float().to(dev)
    q_host = q.float().pin_memory()
    out_host = torch.empty((B, N_CLASS + DOF), dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream(dev)
    hbm_peak, peak_src, sm_max = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(step_fn, steps, warmup, sc=scorer, do_flush=True):
        """Per-step CUDA events summed (max over ranks), launches of this library, clocks, wall seconds."""
        for _ in range(warmup):
            step_fn()
        barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        sampler = ClockSampler(visible_index(local))
        sampler.start()
        l0 = lib.dc_launch_count()
        t0 = time.perf_counter()
        for a, b in evs:
            if do_flush:
                flush.zero_()  # evict the previous step's inputs/outputs from the 126 MB L2 (untimed)
            if world > 1 and sc is not None:
                sc.align()  # untimed: the flushes end at different moments on different GPUs; without this the wait for
                            # the slowest flush would be charged to the step through the step's own collective
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

    def measure(fn, steps, warmup, **kw):
        res = timed(fn, steps, warmup, **kw)
        if any(r in ClockSampler.BAD for r in res[2]["reasons"]):  # rejected: re-measure once
            res = timed(fn, steps, max(3, warmup), **kw)
        return res

    # ---- device-resident throughput (`value`) -------------------------------------------------------------
    ms_dev, launches, clocks, _ = measure(lambda: scorer.score_and_grad(q_dev), args.steps, args.warmup)
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

    # ---- dominant kernel alone (roofline): the fused score+grad launch without the collective; ONE launch per step ----
    ms_k, _, _, _ = timed(lambda: scorer.local_score_and_grad(q_dev), args.steps, 3)
    per_launch_s = ms_k * 1e-3 / args.steps
    which = lib.dc_last_score_kernel()
    kernel_name = {2: "score_tc_kernel<GRAD> (tcgen05 kind::f16, TMEM, bulk TMA; RQ2, C=1, F=14)",
                   1: "score_tq_kernel<F=14,RQ2,C=1,GRAD> (FP32 pipe, bulk TMA)"}.get(which, _lib.KERNEL_NAMES.get(which, "?"))
    kkey = "tensor-core" if which == 2 else "thread-per-query"
    achieved_gbs = BYTES_PER_EVAL * B / per_launch_s / 1e9
    traffic = load_profile_json("roofline_traffic.json").get(kkey, {}).get("dram_bytes_per_launch")
    pipes = load_profile_json("roofline_pipes.json").get(kkey, {})
    sm_mhz = clocks.get("sm_mhz") or sm_max
    fp32_peak_tflops = N_SMS * 128 * 2 * sm_mhz * 1e6 / 1e12
    achieved_tflops = FLOPS_PER_EVAL * B / per_launch_s / 1e12
    pairs_per_clk_sm = B * N_SV / per_launch_s / (N_SMS * sm_mhz * 1e6)
    if which == 2:
        # the binding resource is the warp schedulers' issue slots (4 per SM per clock): the kernel executes
        # `warp_instructions_per_launch` (ncu, cfg-2) in per_launch_s at the clock measured during this run
        wi = pipes.get("warp_instructions_per_launch")
        issue = (wi / (per_launch_s * N_SMS * sm_mhz * 1e6)) if (wi and B == BATCH) else None
        roofline_compute = {"bound": "warp-scheduler issue slots (4 per SM per clock); pipes far from saturated (ncu)",
                            "achieved": issue, "peak": 4.0, "unit": "warp-instructions/clk/SM",
                            "frac": (issue / 4.0 if issue else None), "pairs_per_clk_sm": pairs_per_clk_sm,
                            "sm_mhz_measured": sm_mhz, "ncu": pipes, "algorithmic_tflops": achieved_tflops,
                            "fp32_fma_pipe_peak_tflops": fp32_peak_tflops}
    else:
        roofline_compute = {"bound": "fp32_fma_pipe", "achieved": achieved_tflops, "peak": fp32_peak_tflops,
                            "unit": "TFLOP/s", "frac": achieved_tflops / fp32_peak_tflops,
                            "peak_source": f"{N_SMS} SMs x 128 lanes x 2 flop x {sm_mhz:.0f} MHz (measured during the run)",
                            "algorithmic_flops_per_launch": FLOPS_PER_EVAL * B, "ncu": pipes}

    # ---- end to end through the host-buffer API (`e2e`) -----------------------------------------------------
    ms_e2e, _, _, _ = timed(lambda: scorer.score_and_grad_host(q_host, out_host), args.steps, max(3, args.warmup // 2))
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
    if world > 1:  # this rank's rows of the host result are the device result
        torch.cuda.synchronize(dev)
        s_all, g_all = scorer.score_and_grad(q_dev)
        want = torch.cat([s_all, g_all], dim=1)[rank * B:(rank + 1) * B].cpu()
        scorer.score_and_grad_host(q_host, out_host)
        torch.cuda.synchronize(dev)
        bad = torch.tensor([0 if torch.equal(out_host, want) else 1], device=dev)
        dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        if int(bad.item()) != 0:
            raise SystemExit("bench.py: host-buffer result differs from the device result")

    extras = {}
    if args.extras:
        extras = run_extras(args, torch, dist, lib, dev, rank, world, group, scorer, checker, q_dev, timed, flush, stream, barrier,
                            local, hbm_peak, S, w, q)

    # ---- reference CPU path beside it (rank 0, N == 1 only; bounded sample) -----------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, threads, done, el = cpu_reference_rate(S, w, q, budget_s=10.0)
        rate64, _, done64, el64 = cpu_reference_rate(S, w, q, budget_s=4.0, dtype=torch.float64)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "cpu_model": cpu_model(),
               "sample": f"{done} configurations ({el:.1f} s) of the same workload in {REF_CHUNK}-row chunks, float32, torch CPU",
               "float64": {"value": rate64, "unit": UNIT,
                           "sample": f"{done64} configurations ({el64:.1f} s), float64 (what the reference's scripts run)"}}

    if rank == 0:
        fused_ag = bool(getattr(scorer, "_peer", None))
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_sv": N_SV, "dof": DOF, "n_features": N_FEAT, "batch_per_gpu": B,
                       "global_batch": world * B, "kernel": "RQKernel(gamma=10,p=2)",
                       "sharding": ("none" if world == 1 else
                                    f"batch rows over {world} ranks; all-gather of [score|grad] " +
                                    ("fused into the kernel epilogue (peer stores over NVLink + flag barrier)"
                                     if fused_ag else "by one NCCL all_gather_into_tensor")),
                       "gather_check": gather_check,
                       "arithmetic": ("fp32 in / out; features from a float64 forward kinematics kept as float32 (hi, lo) pairs; the "
                                      "two contractions as tcgen05 kind::f16 MMAs on operands split into two 11-bit terms "
                                      "(products ~22 bits, fp32 accumulation); radial profile and score in fp32; pairs close enough "
                                      "for the tensor core's rho error to matter re-evaluated with direct (hi, lo) differences "
                                      "(tests/test_gpu_tc_stress.py: worst case over 50 seeds x 3 robots x {random, trained})"
                                      if which == 2 else "fp32 throughout (packed FP32 pipe), float64 forward kinematics rounded once"),
                       "l2": "flushed before every step (256 MiB memset, untimed" +
                             ("; ranks re-aligned by an untimed device barrier after it" if world > 1 else "") +
                             "); per-step CUDA events summed"},
            "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved_gbs / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": kernel_name, "kernel_ms": per_launch_s * 1e3,
                         "algorithmic_bytes_per_launch": BYTES_PER_EVAL * B,
                         "note": "the path is compute bound (2600 flop/B), not HBM bound: see roofline_compute"},
            "roofline_compute": roofline_compute,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": q_host.numel() * 4,
                    "d2h_bytes_per_step": out_host.numel() * 4, "ms_per_step": ms_e2e / args.steps,
                    "host_numa": (f"rank 0 pinned to {len(numa_cpus)} GPU-local CPUs before allocating its pinned buffers"
                                  if numa_cpus else "not bound"),
                    "path": ("one launch per rank: the kernel reads the pinned q and writes this rank's records to the pinned "
                             "output (zero-copy over PCIe)" + (" and to every peer's gathered buffer" if world > 1 and fused_ag else "")
                             if which == 2 else "dc_score_grad_host")},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        line.update(extras)
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_extras(args, torch, dist, lib, dev, rank, world, group, scorer, checker, q_dev, timed, flush, stream, barrier, local,
               hbm_peak, S, w, q):
    """Everything beyond the headline line: sustained rate, the autograd surface, the other BASELINE configurations and the
    reference's own ATen code on this GPU.  Every number is device-timed the same way as the headline."""
    from diffco_b200 import DiffCo, MultiDiffCo, _lib
    from diffco_b200 import distributed as D
    from diffco_b200 import kernel as K
    from diffco_b200 import model as M

    out = {}
    B = q_dev.shape[0]

    # ---- sustained: the same launch back to back for >= 2 s (no L2 flush: 1.8 MB in, 2.1 MB out stay L2-resident) --------
    def sustained():
        n, t_dev = 0, 0.0
        sampler = ClockSampler(visible_index(local))
        sampler.start()
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < 2.0:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(200):
                scorer.local_score_and_grad(q_dev)
            b.record(stream)
            torch.cuda.synchronize(dev)
            t_dev += a.elapsed_time(b) * 1e-3
            n += 200
        ck = sampler.stop()
        return n, t_dev, ck

    n_s, t_s, ck_s = sustained()
    out["sustained"] = {"value": world * B * n_s / t_s, "unit": UNIT, "launches": n_s, "seconds": t_s, "us_per_launch": 1e6 * t_s / n_s,
                        "clocks": ck_s, "note": "local kernel, back-to-back launches (no collective, no L2 flush), per rank x ranks"}

    # ---- through the autograd surface the optimisers use: s = checker.score(q); s.sum().backward() ---------------------------
    qa = q_dev.clone().requires_grad_(True)

    def step_autograd():
        qa.grad = None
        checker.score(qa).sum().backward()

    ms_a, _, _, _ = timed(step_autograd, max(5, args.steps // 2), 3, sc=None)
    out["e2e_autograd"] = {"value": world * B * max(5, args.steps // 2) / (ms_a * 1e-3), "unit": UNIT,
                           "note": "DiffCo.score(q).sum().backward() on a device-resident q (autograd.Function around the fused launch)"}

    cfgs = {}
    # ---- cfg5: 2 097 152 configurations split over the ranks (strong scaling), gathered on every rank -----------------------
    total5 = 2097152
    per5 = -(-total5 // world)
    _, _, q5 = make_problem(100 + rank, per5)
    q5 = q5.float().to(dev)
    steps5 = 5
    ms5, _, ck5, _ = timed(lambda: scorer.score_and_grad(q5), steps5, 2)
    t5 = ms5 * 1e-3 / steps5
    cfgs["cfg5"] = {"workload": f"configs[4]: 7-DoF arm, 2000 SVs, {total5} configurations split {per5}/GPU over {world} GPU(s), "
                                "[score|grad] of all ranks gathered on every rank", "scaling": "strong",
                    "value": total5 / t5, "unit": UNIT, "ms_per_step": 1e3 * t5, "steps": steps5,
                    "roofline": {"bound": "hbm", "achieved": BYTES_PER_EVAL * per5 / t5 / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": BYTES_PER_EVAL * per5 / t5 / 1e9 / hbm_peak},
                    "kernel": _lib.KERNEL_NAMES.get(lib.dc_last_score_kernel(), "?")}
    del q5
    if world == 1:
        cfgs["cfg3"] = bench_cfg3(torch, lib, dev, timed, hbm_peak)
        cfgs["wide_tc"] = bench_wide(torch, lib, dev, timed)
        cfgs["fp64"] = bench_fp64(torch, lib, dev, timed, S, w, q)
        cfgs["cfg4"] = bench_cfg4(torch, dev)
        cfgs["fit"] = bench_fit(torch, dev)
        # ---- the reference's own ATen code on this GPU (courtesy row: fused vs unfused on identical silicon) --------------
        try:
            run = _reference_fn(S, w, torch.float32, device=dev)
            run(q_dev)
            torch.cuda.synchronize(dev)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(3):
                run(q_dev)
            b.record(stream)
            torch.cuda.synchronize(dev)
            out["aten_on_b200"] = {"value": 3 * B / (a.elapsed_time(b) * 1e-3), "unit": UNIT,
                                   "note": "the reference algorithm (oracle port: fkine -> cdist -> RQ -> matmul, autograd backward) "
                                           "on CUDA tensors, float32, full 65536-row batch: ATen kernels, B x N intermediates in HBM"}
        except Exception as e:  # e.g. out of memory on a shared device
            out["aten_on_b200"] = {"unavailable": str(e)[:120]}
    out["configs"] = cfgs
    return out


def bench_cfg3(torch, lib, dev, timed, hbm_peak):
    """BASELINE configs[2]: Baxter left arm (F = 12), legacy MultiDiffCo with 4 classes, 5000 SVs, rbf_score with
    Polyharmonic(1, 1), batch 262144, score + gradient of the class sum."""
    from diffco_b200 import MultiDiffCo, _lib
    from diffco_b200 import distributed as D
    from diffco_b200 import kernel as K
    from diffco_b200 import model as M

    n, c, b = 5000, 4, 262144
    robot = M.BaxterLeftArmFK()
    gen = torch.Generator().manual_seed(SEED)
    lim = robot.limits.double()
    Sc = torch.rand(n, 7, generator=gen, dtype=torch.float64) * (lim[:, 1] - lim[:, 0]) + lim[:, 0]
    W = torch.randn(n, c, generator=gen, dtype=torch.float64) * (torch.rand(n, c, generator=gen) < 0.6)
    qc = torch.rand(b, 7, generator=gen, dtype=torch.float64) * (lim[:, 1] - lim[:, 0]) + lim[:, 0]
    chk = MultiDiffCo(None, kernel_func=K.RQKernel(GAMMA), transform=robot.fkine)
    chk.num_class = c
    chk.support_points = Sc.float().to(dev)
    chk.support_transformed = robot.fkine(chk.support_points)
    chk.gains = W.float().to(dev)
    chk.rbf_nodes, chk.rbf_kernel = W.float().to(dev), K.Polyharmonic(1, 1.0)
    sc = D.ShardedScorer(chk, weights="rbf")
    qd = qc.float().to(dev)
    steps = 10
    ms, _, _, _ = timed(lambda: sc.local_score_and_grad(qd), steps, 3, sc=None)
    t = ms * 1e-3 / steps
    bytes_per_eval = 4 * (2 * 7 + 2 * c)
    flops = n * (5 * 12 + 8 + 4 * c)
    return {"workload": "configs[2]: Baxter 7-DoF (F=12), MultiDiffCo 4 classes, 5000 SVs, Polyharmonic(1,1) rbf_score, batch 262144, "
                        "score + gradient of the class sum, fp32", "value": b / t, "unit": UNIT, "ms_per_step": 1e3 * t, "steps": steps,
            "kernel": _lib.KERNEL_NAMES.get(lib.dc_last_score_kernel(), "?") + " (score_tq_kernel<F=12,PH1,C=4,GRAD>: FP32 pipe)",
            "roofline": {"bound": "hbm", "achieved": bytes_per_eval * b / t / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": bytes_per_eval * b / t / 1e9 / hbm_peak, "algorithmic_bytes_per_eval": bytes_per_eval},
            "algorithmic_tflops": flops * b / t / 1e12}


def bench_wide(torch, lib, dev, timed):
    """The tensor-core kernel's wide instantiation (15 <= F <= 30, one CTA per SM) on the reference tutorial's robot:
    PandaFK (F = 21), 2000 SVs, 65536 queries, RQKernel — at the reference's default width (gamma = 10: metre-scale
    features make the kernel wide, the dispatcher decides) and at a narrow width (gamma = 300), each against the FP32-pipe
    kernels (tensor-core path switched off)."""
    from diffco_b200 import DiffCo, _lib
    from diffco_b200 import distributed as D
    from diffco_b200 import kernel as K
    from diffco_b200 import model as M

    n, b = 2000, 65536
    robot = M.PandaFK()
    gen = torch.Generator().manual_seed(SEED)
    lim = robot.limits.double()
    Sc = torch.rand(n, 7, generator=gen, dtype=torch.float64) * (lim[:, 1] - lim[:, 0]) + lim[:, 0]
    W = torch.randn(n, generator=gen, dtype=torch.float64)
    qd = (torch.rand(b, 7, generator=gen, dtype=torch.float64) * (lim[:, 1] - lim[:, 0]) + lim[:, 0]).float().to(dev)
    out = {"workload": "PandaFK (F = 21), 2000 SVs, batch 65536, RQKernel, score + gradient, fp32", "unit": UNIT}
    tc_was = lib.dc_get_option(_lib.DC_OPT_TC_ENABLE)
    try:
        for gamma in (10.0, 300.0):
            chk = DiffCo(kernel_func=K.RQKernel(gamma), transform=robot.fkine)
            chk.support_points = Sc.float().to(dev)
            chk.support_transformed = robot.fkine(chk.support_points)
            chk.gains = W.float().to(dev)
            sc = D.ShardedScorer(chk, weights="gains")
            row = {}
            for label, tc in (("dispatch", 1.0), ("fp32_pipe", 0.0)):
                _lib.check(lib.dc_set_option(_lib.DC_OPT_TC_ENABLE, tc), "dc_set_option")
                ms, _, _, _ = timed(lambda: sc.local_score_and_grad(qd), 10, 3, sc=None)
                row[label] = {"value": b * 10 / (ms * 1e-3), "kernel": _lib.KERNEL_NAMES.get(lib.dc_last_score_kernel(), "?")}
            out[f"gamma_{gamma:g}"] = row
    finally:
        lib.dc_set_option(_lib.DC_OPT_TC_ENABLE, tc_was)
    return out


def bench_fp64(torch, lib, dev, timed, S, w, q):
    """The headline workload in float64 — the dtype the reference's scripts run in (its CPU figure: cpu_baseline.float64):
    same supports and queries, lane-split kernel (the float64 path is built for the optimisers' small batches; this is its
    large-batch rate)."""
    from diffco_b200 import DiffCo, _lib
    from diffco_b200 import distributed as D
    from diffco_b200 import kernel as K
    from diffco_b200 import model as M

    robot = M.RevolutePlanarRobot(1.0, 0.3, dof=DOF)
    chk = DiffCo(kernel_func=K.RQKernel(GAMMA), transform=robot.fkine)
    chk.support_points = S.double().to(dev)
    chk.support_transformed = robot.fkine(chk.support_points)
    chk.gains = w.double().to(dev)
    sc = D.ShardedScorer(chk, weights="gains")
    qd = q.double().to(dev)
    steps = 5
    ms, _, _, _ = timed(lambda: sc.local_score_and_grad(qd), steps, 2, sc=None)
    t = ms * 1e-3 / steps
    return {"workload": WORKLOAD.replace("fp32", "fp64"), "value": len(q) / t, "unit": UNIT, "ms_per_step": 1e3 * t, "steps": steps,
            "kernel": _lib.KERNEL_NAMES.get(lib.dc_last_score_kernel(), "?") + " (score_ls_kernel<double>)"}


def bench_cfg4(torch, dev):
    """BASELINE configs[3]: SE(2) base + 3-link arm, 3000 SVs, 256 waypoints, Weighted.step Adam loop ('MultiFourier' does not
    exist in the reference: MultiQuadratic stands in) — microseconds per graphed step vs the autograd step."""
    from diffco_b200 import DiffCo, trajopt
    from diffco_b200 import kernel as K
    from diffco_b200 import model as M
    from diffco_b200 import optim as OPT

    robot = M.SE2BasePlanarArm([[0.5, -0.5, -0.5, 0.5], [0.3, 0.3, -0.3, -0.3]], [1.0, 1.0, 1.0])
    gen = torch.Generator().manual_seed(SEED)
    lim = robot.limits.double()
    Sc = torch.rand(3000, robot.dof, generator=gen, dtype=torch.float64) * (lim[:, 1] - lim[:, 0]) + lim[:, 0]
    W = torch.randn(3000, generator=gen, dtype=torch.float64)
    dc = DiffCo(kernel_func=K.RQKernel(GAMMA), transform=robot.fkine)
    dc.support_points = Sc.float().to(dev)
    dc.support_transformed = robot.fkine(dc.support_points)
    dc.gains = W.float().to(dev)
    dc.rbf_nodes, dc.rbf_kernel = 0.05 * W.float().to(dev), K.MultiQuadratic(1.0)
    ab = torch.rand(2, robot.dof, generator=gen, dtype=torch.float64) * (lim[:, 1] - lim[:, 0]) + lim[:, 0]
    init = (ab[0] + (ab[1] - ab[0]) * torch.linspace(0, 1, 256, dtype=torch.float64)[:, None]).float()
    steps = 200
    options = {"n_waypoints": 256, "maxiter": steps, "history": False, "max_move_weight": 10, "collision_weight": 10,
               "joint_limit_weight": 10, "safety_bias": 1e6, "max_speed": 0.3, "optimizer": torch.optim.Adam,
               "optimizer_params": {"lr": 0.02}, "dense_check": False}  # huge bias: the early exit never triggers
    mask = torch.ones(256, dtype=torch.bool)
    mask[[0, -1]] = False
    wf = OPT.Weighted(robot, dc, dict(options, fused=True))
    wf.step(init.clone(), maxiter=3, mask=mask)  # builds (and caches) the graph
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    wf.step(init.clone(), mask=mask)
    torch.cuda.synchronize(dev)
    t_graph = (time.perf_counter() - t0) / steps
    wa = OPT.Weighted(robot, dc, dict(options))
    wa.step(init.clone(), maxiter=3, mask=mask)
    t0 = time.perf_counter()
    wa.step(init.clone(), maxiter=50, mask=mask)
    torch.cuda.synchronize(dev)
    t_auto = (time.perf_counter() - t0) / 50
    return {"workload": "configs[3]: SE(2) base + 3-link arm, MultiQuadratic (stands in for 'MultiFourier'), 3000 SVs, 256 waypoints, "
                        "Weighted.step Adam loop", "value": 1e6 * t_graph, "unit": "us/step (CUDA-graph replay: dc_score_grad + dc_traj_step)",
            "higher_is_better": False, "steps": steps, "autograd_step_us": 1e6 * t_auto,
            "note": "wall clock over the whole Weighted.step call (host loop included); second call, graph cached"}


def bench_fit(torch, dev):
    """ForwardKinematicsDiffCo.fit at the size of the reference's tutorial (Panda 7-DoF, 10 000 samples, RQ gamma = 10;
    published there: 848 iterations in 0.381 s = 2256 it/s on the author's workstation).  The ground truth is
    synthetic (control points inside spheres); the perceptron trains in ONE persistent launch (dc_perceptron_train_rows)."""
    from diffco_b200 import ForwardKinematicsDiffCo
    from diffco_b200 import model as M

    robot = M.PandaFK()
    centres = torch.tensor([[0.45, 0.0, 0.55], [-0.1, 0.45, 0.4], [0.2, -0.4, 0.8]], dtype=torch.float32, device=dev)
    radii = torch.tensor([0.22, 0.18, 0.2], dtype=torch.float32, device=dev)

    def gt(q):
        pts = robot.fkine(q.to(dev).float())  # (B, M, 3)
        hit = ((pts[:, :, None, :] - centres[None, None]).norm(dim=-1) < radii).any(dim=2).any(dim=1)
        return hit.to(q.dtype)

    chk = ForwardKinematicsDiffCo(robot=robot, gt_check_func=gt, device=dev)
    gen = torch.Generator().manual_seed(SEED)
    lim = robot.limits.double()
    X = (torch.rand(10000, 7, generator=gen, dtype=torch.float64) * (lim[:, 1] - lim[:, 0]) + lim[:, 0]).float().to(dev)
    labels = gt(X)
    chk.perceptron.train(X[:512], 2 * labels[:512] - 1, max_iteration=512)  # warm-up (module load, allocator)
    torch.cuda.synchronize(dev)
    y = 2 * labels - 1
    t0 = time.perf_counter()
    chk.perceptron.train(X, y, max_iteration=len(X))
    torch.cuda.synchronize(dev)
    t_train = time.perf_counter() - t0
    iters = chk.perceptron.train_iterations + 1
    t0 = time.perf_counter()
    rates = chk.fit(q=X, labels=labels, verify_ratio=0.1)
    torch.cuda.synchronize(dev)
    t_fit = time.perf_counter() - t0
    return {"workload": "ForwardKinematicsDiffCo.fit: PandaFK (F = 21), 10000 samples (synthetic sphere obstacles, "
                        f"{float(labels.mean()):.2f} in collision), RQKernel(10), fp32", "value": iters / t_train,
            "unit": "training iterations/s (train_perceptron only)", "iterations": iters, "train_s": t_train,
            "support_points": int(len(chk.perceptron.support_points)), "kernel_rows_computed": chk.perceptron.train_kernel_rows,
            "fit_s": t_fit, "fit_note": "fit() = split + train + fit_poly(Polyharmonic(1,1)) + safety bias + verification",
            "verify_acc_tpr_tnr": [float(v) for v in rates],
            "published_reference": "848 it, 0.381 s loop (2256 it/s), 467 supports, author workstation CPU (BASELINE.md §1)"}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


if __name__ == "__main__":
    sys.exit(main())
