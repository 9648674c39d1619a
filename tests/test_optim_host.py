"""CPU tests of the optimiser drivers (diffco_b200/optim.py, utils.py): host logic only.  dist_est and the robot are
oracle-backed stubs here (the CUDA dist_est is exercised by tests/test_gpu_parity.py / test_gpu_optim.py); the records
must reproduce what the UNMODIFIED reference optimisers produced (tests/golden/optim_replay.npz) and, when
/root/reference is present, the live reference run side by side."""
import os
import types
import warnings

import numpy as np
import pytest
import torch

from diffco_b200 import optim as OPT
from diffco_b200 import utils as U
from oracle import diffco_oracle as O
from oracle import ref_loader

GOLD = os.path.join(os.path.dirname(__file__), "golden")
T64 = lambda a: torch.from_numpy(np.asarray(a)).double()


def golden_problem():
    g = np.load(os.path.join(GOLD, "optim_replay.npz"))
    L = torch.ones(7, dtype=torch.float64)
    fk = lambda q: O.fk_planar_chain(q, L)
    limits = torch.FloatTensor([[-np.pi, np.pi]] * 7)
    robot = types.SimpleNamespace(dof=7, limits=limits, fkine=fk, wrap=U.wrap2pi)
    St, nodes = fk(T64(g["support_points"])), T64(g["nodes"])
    ph = O.KernelSpec("polyharmonic", 1.0, 1)
    dist_est = lambda p: O.poly_score(p, fk, ph, St, nodes)
    start = torch.tensor([-2.0, -0.4, 0.3, -0.2, 0.1, 0.2, -0.1], dtype=torch.float64)
    target = torch.tensor([1.6, 0.5, -0.3, 0.4, -0.2, 0.1, 0.3], dtype=torch.float64)
    init = torch.from_numpy(np.linspace(start.numpy(), target.numpy(), 12))
    opts = {"N_WAYPOINTS": 12, "NUM_RE_TRIALS": 1, "MAXITER": 15, "safety_margin": -0.3, "max_speed": 0.6, "seed": 1234,
            "history": False, "extra_optimizer_options": {"lr": 0.05}, "init_solution": init.clone()}
    return g, robot, dist_est, start, target, init, opts


def test_dense_path_matches_reference_semantics():
    q = torch.tensor([[0.0, 0.0], [1.0, 0.0], [1.0, 2.5], [1.0, 2.5001]], dtype=torch.float64, requires_grad=True)
    d = U.dense_path(q, max_step=0.4)
    ref = O.dense_path(q.detach(), 0.4)
    assert d.shape == ref.shape and torch.allclose(d, ref, atol=1e-15)
    (d**2).sum().backward()  # differentiable through the normalised direction
    qr = q.detach().clone().requires_grad_(True)
    (O.dense_path(qr, 0.4) ** 2).sum().backward()
    assert torch.allclose(q.grad, qr.grad, atol=1e-12)
    d2 = U.dense_path(q.detach(), max_step=0.1, max_step_num=6)  # step size raised so that at most ~6 points are used
    assert torch.allclose(d2, O.dense_path(q.detach(), 0.1, 6), atol=1e-15)
    assert torch.allclose(U.wrap2pi(torch.tensor([3.5, -3.5, 0.0])), torch.tensor([3.5 - 2 * np.pi, 2 * np.pi - 3.5, 0.0]))


def test_adam_traj_optimize_reproduces_reference_record():
    g, robot, dist_est, start, target, init, opts = golden_problem()
    rec = OPT.adam_traj_optimize(robot, dist_est, start, target, dict(opts))
    assert set(rec) == {"start_cfg", "target_cfg", "cnt_check", "cost", "time", "success", "seed", "solution"}
    assert np.abs(np.array(rec["solution"]) - g["adam_solution"]).max() <= 1e-8
    assert abs(rec["cost"] - float(g["adam_cost"])) <= 1e-8 * max(1.0, abs(float(g["adam_cost"])))
    assert rec["cnt_check"] == 12 * int(g["n_calls"][0])


def test_givengrad_traj_optimize_reproduces_reference_record():
    g, robot, dist_est, start, target, init, opts = golden_problem()
    opts = dict(opts, MAXITER=6, extra_optimizer_options={"ftol": 1e-4, "disp": False}, init_solution=init.clone())
    rec = OPT.givengrad_traj_optimize(robot, dist_est, start, target, opts)
    assert np.abs(np.array(rec["solution"]) - g["slsqp_solution"]).max() <= 1e-6
    assert abs(rec["cost"] - float(g["slsqp_cost"])) <= 1e-6 * max(1.0, abs(float(g["slsqp_cost"])))


def test_two_point_initial_solution_short_circuits():
    g, robot, dist_est, start, target, init, opts = golden_problem()
    opts = dict(opts, init_solution=torch.stack([start, target]))
    for fn in (OPT.adam_traj_optimize, OPT.givengrad_traj_optimize):
        rec = fn(robot, dist_est, start, target, dict(opts))
        assert rec["success"] and rec["cnt_check"] == 0 and len(rec["solution"]) == 2


def test_unsupported_optimisers_fail_loudly():
    with pytest.raises(NotImplementedError):
        OPT.gradient_free_traj_optimize(None, None, None, None, {})


def test_constraint_hessian_matches_double_backward():
    """optim.py:380-391: the reference gets sum_k v_k Hessian(c_k) by a double backward through its autograd kernel; ours
    is a central difference of the analytic first derivative, all perturbed paths in one dist_est call.  Same matrix."""
    g, robot, dist_est, start, target, init, opts = golden_problem()
    path = init.clone()
    path[1:-1] += 0.3 * torch.randn(10, 7, generator=torch.Generator().manual_seed(3), dtype=torch.float64)
    prob = OPT._SlsqpProblem(robot, dist_est, path, -0.3, 0.6)
    x = path[1:-1].reshape(-1).numpy()
    assert (prob.con_collision(x) < 0).any(), "the test path must violate the constraint somewhere"
    v = torch.randn(11, generator=torch.Generator().manual_seed(4), dtype=torch.float64)
    got = prob.hess_con_collision(x, v.numpy(), fd_step=1e-5)
    full = prob.full_path(x).detach()
    want = torch.autograd.functional.hessian(lambda p: torch.dot(prob.collision_tensor(p), v), full, vectorize=False)
    want = want[1:-1, :, 1:-1, :].reshape(70, 70).numpy()
    assert got.shape == (70, 70) and np.abs(want).max() > 0
    assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max()
    assert np.allclose(prob.hess_con_collision(x, np.zeros(11)), 0)


def test_trustconstr_traj_optimize_runs_and_reduces_violation():
    g, robot, dist_est, start, target, init, opts = golden_problem()
    opts = dict(opts, MAXITER=12, extra_optimizer_options={"verbose": 0}, init_solution=init.clone(), hess_fd_step=1e-5)
    rec = OPT.trustconstr_traj_optimize(robot, dist_est, start, target, opts)
    assert set(rec) == {"start_cfg", "target_cfg", "cnt_check", "cost", "time", "success", "seed", "solution", "info"}
    sol = torch.tensor(rec["solution"], dtype=torch.float64)
    assert sol.shape == init.shape and torch.equal(sol[0], start) and torch.equal(sol[-1], target) and rec["cnt_check"] > 0
    prob = OPT._SlsqpProblem(robot, dist_est, init, -0.3, 0.6)
    before = -prob.con_collision(init[1:-1].reshape(-1).numpy()).sum()
    after = -prob.con_collision(sol[1:-1].reshape(-1).numpy()).sum()
    assert after <= before


def test_weighted_step_autograd_path_with_stub_checker():
    g, robot, dist_est, start, target, init, opts = golden_problem()
    checker = types.SimpleNamespace(device=torch.device("cpu"), rbf_score=dist_est)
    options = {"n_waypoints": 12, "maxiter": 8, "history": True, "max_move_weight": 10, "collision_weight": 10,
               "joint_limit_weight": 10, "safety_bias": 0.3, "max_speed": 0.6, "optimizer": torch.optim.Adam,
               "optimizer_params": {"lr": 0.05}, "dense_check": True}
    w = OPT.Weighted(robot, checker, options)
    res = w.step(init.clone())
    assert isinstance(res, OPT.OptimizerResult) and res.x.shape == init.shape and len(res.misc["path_history"]) <= 8
    # same loop written out with the oracle pieces
    p = init.clone().requires_grad_(True)
    opt = torch.optim.Adam([p], lr=0.05)
    for _ in range(len(res.misc["path_history"])):
        opt.zero_grad()
        col = torch.clamp(dist_est(O.dense_path(p, 0.6)) + 0.3, min=0).mean() * len(p)
        cp = robot.fkine(p)
        seg = (cp[1:] - cp[:-1]).square()
        mm = torch.clamp(seg.sum(dim=2) - 0.36, min=0).sum()
        jl = (torch.clamp(robot.limits[:, 0] - p, min=0) + torch.clamp(p - robot.limits[:, 1], min=0)).sum()
        (seg.sum() + 10 * col + 10 * mm + 10 * jl).backward()
        opt.step()
        p.data = U.wrap2pi(p.data)
    assert torch.allclose(res.x, p.detach(), atol=1e-9)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
def test_optimisers_against_live_reference():
    warnings.filterwarnings("ignore")
    ns = ref_loader.load()
    g, robot, dist_est, start, target, init, opts = golden_problem()
    for steps, lr in ((25, 0.1), (10, 0.02)):
        o = dict(opts, MAXITER=steps, extra_optimizer_options={"lr": lr}, init_solution=init.clone())
        ours = OPT.adam_traj_optimize(robot, dist_est, start, target, dict(o))
        ref = ns.optim.adam_traj_optimize(robot, dist_est, start, target, dict(o, init_solution=init.clone()))
        assert np.abs(np.array(ours["solution"]) - np.array(ref["solution"])).max() <= 1e-9
        assert ours["success"] == ref["success"] and ours["cnt_check"] == ref["cnt_check"]
    o = dict(opts, MAXITER=8, extra_optimizer_options={"ftol": 1e-5, "disp": False}, init_solution=init.clone())
    ours = OPT.givengrad_traj_optimize(robot, dist_est, start, target, dict(o))
    ref = ns.optim.givengrad_traj_optimize(robot, dist_est, start, target, dict(o, init_solution=init.clone()))
    assert np.abs(np.array(ours["solution"]) - np.array(ref["solution"])).max() <= 1e-6
    assert ours["success"] == ref["success"]
