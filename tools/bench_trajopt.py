#!/usr/bin/env python
"""Per-step time of the Weighted.step Adam loop at BASELINE.json configs[3] (SE(2) base + 3-link arm, 3000 SVs,
256 waypoints): CUDA-graph-replayed analytic step (diffco_b200/trajopt.py) vs the autograd step on the same GPU kernels
vs the reference algorithm on the host cores (oracle port, float64 as the reference's scripts run it).  Prints one JSON line.

    python tools/bench_trajopt.py [--steps 200]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from tests import problems as P  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    args = ap.parse_args()
    from diffco_b200 import DiffCo, trajopt
    from diffco_b200 import kernel as K
    from diffco_b200 import optim as OPT
    from oracle import diffco_oracle as O

    dev = torch.device("cuda", 0)
    robot, S, W = P.synthetic_model("se2arm", 3000, 1, seed=1234)
    dc = DiffCo(kernel_func=K.RQKernel(10.0), transform=robot.fkine)
    dc.support_points = S.float().to(dev)
    dc.support_transformed = robot.fkine(dc.support_points)
    dc.gains = W[:, 0].float().to(dev)
    dc.rbf_nodes, dc.rbf_kernel = 0.05 * W[:, 0].float().to(dev), K.MultiQuadratic(1.0)
    gen = torch.Generator().manual_seed(7)
    a, b = P.sample_configs(robot, 2, gen).float()
    init = a + (b - a) * torch.linspace(0, 1, 256)[:, None]
    options = {"n_waypoints": 256, "maxiter": args.steps, "history": False, "max_move_weight": 10, "collision_weight": 10,
               "joint_limit_weight": 10, "safety_bias": 1e6, "max_speed": 0.3, "optimizer": torch.optim.Adam,
               "optimizer_params": {"lr": 0.02}, "dense_check": False}  # huge bias: the early exit never triggers
    mask = torch.ones(256, dtype=torch.bool)
    mask[[0, -1]] = False
    w = OPT.Weighted(robot, dc, dict(options, fused=True))
    t0 = time.perf_counter()
    g = trajopt.GraphedWeightedStep(w, init.to(dev), mask)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    t0 = time.perf_counter()
    g.run(args.steps)
    torch.cuda.synchronize()
    t_graph = (time.perf_counter() - t0) / args.steps
    wa = OPT.Weighted(robot, dc, dict(options))
    wa.step(init.clone(), maxiter=3, mask=mask)
    t0 = time.perf_counter()
    wa.step(init.clone(), mask=mask)
    torch.cuda.synchronize()
    t_auto = (time.perf_counter() - t0) / args.steps
    # reference algorithm on the host cores: poly_score + autograd through FK / cdist-free MultiQuadratic (oracle port)
    torch.set_num_threads(os.cpu_count() or 1)
    fk = P.oracle_fk(robot)
    St = fk(S).reshape(len(S), -1)
    nodes = 0.05 * W[:, 0]
    mq = O.KernelSpec("multiquadric", 1.0, 0)
    lim = robot.limits.double()
    p = init.double().clone().requires_grad_(True)
    opt = torch.optim.Adam([p], lr=0.02)
    n_cpu = max(5, min(20, args.steps))
    t0 = time.perf_counter()
    for _ in range(n_cpu):
        opt.zero_grad()
        col = torch.clamp(O.poly_score(p, fk, mq, St, nodes) + 1e6, min=0).mean() * len(p)
        cp = fk(p)
        seg = (cp[1:] - cp[:-1]).square()
        mm = torch.clamp(seg.sum(dim=2) - 0.09, min=0).sum()
        jl = (torch.clamp(lim[:, 0] - p, min=0) + torch.clamp(p - lim[:, 1], min=0)).sum()
        (seg.sum() + 10 * col + 10 * mm + 10 * jl).backward()
        p.grad[~mask] = 0
        opt.step()
    t_cpu = (time.perf_counter() - t0) / n_cpu
    print(json.dumps({"workload": "configs[3]: SE(2) base + 3-link arm, MultiQuadratic (stands in for 'MultiFourier'), 3000 SVs, "
                                  "256 waypoints, Weighted.step Adam loop", "steps": args.steps,
                      "graphed_step_us": 1e6 * t_graph, "graph_build_ms": 1e3 * t_build, "autograd_step_us": 1e6 * t_auto,
                      "reference_cpu_step_us": 1e6 * t_cpu, "cpu_threads": os.cpu_count(),
                      "speedup_vs_autograd": t_auto / t_graph, "speedup_vs_reference_cpu": t_cpu / t_graph}))


if __name__ == "__main__":
    main()
