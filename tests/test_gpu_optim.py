"""GPU tests: the reference's trajectory-optimisation call stacks end to end on the CUDA path — diffco_b200.optim drivers,
diffco_b200 robot (device FK with VJP) and diffco_b200 DiffCo (fused score kernel + analytic Jacobian) — against the
records the UNMODIFIED reference produced for the same problem (tests/golden/optim_replay.npz).  Gate:
same success flag, solution within 1e-4."""
import os

import numpy as np
import pytest
import torch

from tests import problems as P

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
T64 = lambda a: torch.from_numpy(np.asarray(a)).double()


@pytest.fixture(scope="module")
def problem(cuda_device):
    from diffco_b200 import DiffCo
    from diffco_b200 import kernel as K

    g = np.load(os.path.join(GOLD, "optim_replay.npz"))
    robot = P.make_robot("planar7")
    dc = DiffCo(kernel_func=K.RQKernel(10.0), transform=robot.fkine, beta=1.0)
    dc.train(T64(g["X"]), T64(g["y"]), max_iteration=len(g["X"]))
    assert dc.support_index.tolist() == g["idx"].tolist()
    dc.fit_poly(K.Polyharmonic(1, 1.0), target="label")
    start = torch.tensor([-2.0, -0.4, 0.3, -0.2, 0.1, 0.2, -0.1], dtype=torch.float64)
    target = torch.tensor([1.6, 0.5, -0.3, 0.4, -0.2, 0.1, 0.3], dtype=torch.float64)
    init = torch.from_numpy(np.linspace(start.numpy(), target.numpy(), 12))
    opts = {"N_WAYPOINTS": 12, "NUM_RE_TRIALS": 1, "MAXITER": 15, "safety_margin": -0.3, "max_speed": 0.6, "seed": 1234,
            "history": False, "extra_optimizer_options": {"lr": 0.05}, "init_solution": init.clone()}
    return g, robot, dc, start, target, init, opts


def test_adam_traj_optimize_on_cuda_matches_reference(problem):
    from diffco_b200 import optim as OPT

    g, robot, dc, start, target, init, opts = problem
    rec = OPT.adam_traj_optimize(robot, dc.poly_score, start, target, dict(opts, init_solution=init.clone()))
    assert np.abs(np.array(rec["solution"]) - g["adam_solution"]).max() <= 1e-4
    assert abs(rec["cost"] - float(g["adam_cost"])) <= 1e-4 * max(1.0, abs(float(g["adam_cost"])))


def test_givengrad_traj_optimize_on_cuda_matches_reference(problem):
    from diffco_b200 import optim as OPT

    g, robot, dc, start, target, init, opts = problem
    o = dict(opts, MAXITER=6, extra_optimizer_options={"ftol": 1e-4, "disp": False}, init_solution=init.clone())
    rec = OPT.givengrad_traj_optimize(robot, dc.poly_score, start, target, o)
    assert np.abs(np.array(rec["solution"]) - g["slsqp_solution"]).max() <= 1e-4
    assert abs(rec["cost"] - float(g["slsqp_cost"])) <= 1e-4 * max(1.0, abs(float(g["slsqp_cost"])))


def test_trustconstr_on_cuda_hessian_and_record(problem):
    """optim.py:324-507 on the CUDA dist_est: the finite-difference constraint Hessian (one fused launch over all perturbed
    paths) against the double backward of the float64 oracle on the same model, then a short trust-constr run."""
    from diffco_b200 import optim as OPT
    from oracle import diffco_oracle as O

    g, robot, dc, start, target, init, opts = problem
    path = init.clone()
    path[1:-1] += 0.3 * torch.randn(10, 7, generator=torch.Generator().manual_seed(3), dtype=torch.float64)
    x = path[1:-1].reshape(-1).numpy()
    prob = OPT._SlsqpProblem(robot, dc.poly_score, path, -0.3, 0.6)
    v = torch.randn(11, generator=torch.Generator().manual_seed(4), dtype=torch.float64)
    got = prob.hess_con_collision(x, v.numpy())
    fk = lambda q: O.fk_planar_chain(q, torch.ones(7, dtype=torch.float64))
    St, nodes = fk(dc.support_points.double().cpu()), dc.rbf_nodes.double().cpu().reshape(-1)
    ref_est = lambda p: O.poly_score(p, fk, O.KernelSpec("polyharmonic", 1.0, 1), St, nodes)
    ref = OPT._SlsqpProblem(robot, ref_est, path, -0.3, 0.6)
    want = torch.autograd.functional.hessian(lambda p: torch.dot(ref.collision_tensor(p), v), ref.full_path(x).detach())
    want = want[1:-1, :, 1:-1, :].reshape(70, 70).numpy()
    assert np.abs(want).max() > 0
    assert np.abs(got - want).max() <= 2e-2 * np.abs(want).max(), np.abs(got - want).max() / np.abs(want).max()
    o = dict(opts, MAXITER=10, extra_optimizer_options={"verbose": 0}, init_solution=init.clone())
    rec = OPT.trustconstr_traj_optimize(robot, dc.poly_score, start, target, o)
    sol = torch.tensor(rec["solution"], dtype=torch.float64)
    assert sol.shape == init.shape and torch.equal(sol[0], start) and torch.equal(sol[-1], target)
    before = -prob.con_collision(init[1:-1].reshape(-1).numpy()).sum()
    after = -prob.con_collision(sol[1:-1].reshape(-1).numpy()).sum()
    assert after <= before + 1e-9 and rec["cnt_check"] > 0


def test_weighted_step_device_resident(problem):
    """Weighted.step (optim.py:686-761) with the waypoints on the GPU: autograd path vs the same loop on the oracle."""
    from diffco_b200 import optim as OPT
    from diffco_b200 import utils as U
    from oracle import diffco_oracle as O

    g, robot, dc, start, target, init, opts = problem
    options = {"n_waypoints": 12, "maxiter": 6, "history": False, "max_move_weight": 10, "collision_weight": 10,
               "joint_limit_weight": 10, "safety_bias": 0.3, "max_speed": 0.6, "optimizer": torch.optim.Adam,
               "optimizer_params": {"lr": 0.05}, "dense_check": True}
    res = OPT.Weighted(robot, dc, options).step(init.clone())
    fk = P.oracle_fk(robot)
    St, nodes = fk(T64(g["support_points"])), T64(g["nodes"])
    ph = O.KernelSpec("polyharmonic", 1.0, 1)
    p = init.clone().requires_grad_(True)
    opt = torch.optim.Adam([p], lr=0.05)
    lim = robot.limits.double()
    for _ in range(6):
        opt.zero_grad()
        col = torch.clamp(O.poly_score(O.dense_path(p, 0.6), fk, ph, St, nodes) + 0.3, min=0).mean() * len(p)
        cp = fk(p)
        seg = (cp[1:] - cp[:-1]).square()
        mm = torch.clamp(seg.sum(dim=2) - 0.36, min=0).sum()
        jl = (torch.clamp(lim[:, 0] - p, min=0) + torch.clamp(p - lim[:, 1], min=0)).sum()
        loss = seg.sum() + 10 * col + 10 * mm + 10 * jl
        loss.backward()
        opt.step()
        p.data = U.wrap2pi(p.data)
        if float((10 * col + 10 * mm + 10 * jl).detach()) <= 0.5:
            break
    assert P.rel_to_max(res.x, p.detach()) <= 1e-5


def test_weighted_step_graphed_matches_autograd_path(problem):
    """options['fused'] = True (diffco_b200/trajopt.py): the analytic, CUDA-graph-replayed step must walk the same
    trajectory as the autograd step of the reference's Weighted.step (optim.py:686-761) — float64, 25 Adam steps."""
    from diffco_b200 import optim as OPT

    g, robot, dc, start, target, init, opts = problem
    options = {"n_waypoints": 12, "maxiter": 25, "history": True, "max_move_weight": 10, "collision_weight": 10,
               "joint_limit_weight": 10, "safety_bias": 0.3, "max_speed": 0.6, "optimizer": torch.optim.Adam,
               "optimizer_params": {"lr": 0.05}, "dense_check": False}
    mask = torch.ones(12, dtype=torch.bool)
    mask[[0, -1]] = False
    ref = OPT.Weighted(robot, dc, dict(options)).step(init.clone(), mask=mask)
    fused = OPT.Weighted(robot, dc, dict(options, fused=True)).step(init.clone(), mask=mask)
    assert len(fused.misc["path_history"]) == len(ref.misc["path_history"])
    assert (fused.x[[0, -1]] - init[[0, -1]]).abs().max() <= 1e-12  # end points are masked (wrap() may round them)
    assert P.rel_to_max(fused.x, ref.x) <= 1e-9
    for a, b in zip(fused.misc["path_history"], ref.misc["path_history"]):
        assert P.rel_to_max(a, b) <= 1e-9
    # dense collision checking on the graphed step (device-side dense_path) against the autograd step with utils.dense_path
    ref_d = OPT.Weighted(robot, dc, dict(options, dense_check=True)).step(init.clone(), mask=mask)
    opt_d = OPT.Weighted(robot, dc, dict(options, fused=True, dense_check=True))
    fused_d = opt_d.step(init.clone(), mask=mask)
    assert len(fused_d.misc["path_history"]) == len(ref_d.misc["path_history"])
    assert P.rel_to_max(fused_d.x, ref_d.x) <= 1e-9
    # a second call reuses the captured graph (no re-capture) and gives the same answer; without history the exit state is
    # read back only every few iterations and must stop at exactly the same iterate
    stepper = opt_d._fused_cache[1]
    again = opt_d.step(init.clone(), mask=mask)
    assert opt_d._fused_cache[1] is stepper and torch.equal(again.x, fused_d.x)
    lazy = OPT.Weighted(robot, dc, dict(options, fused=True, dense_check=True, history=False)).step(init.clone(), mask=mask)
    assert torch.equal(lazy.x, fused_d.x)
    with pytest.raises(RuntimeError):  # more interpolation points than the static bound
        OPT.Weighted(robot, dc, dict(options, fused=True, dense_check=True, max_dense_points=13, max_speed=0.01)).step(init.clone())


def test_weighted_step_graphed_cfg4_256_waypoints():
    """BASELINE.json configs[3]: SE(2) base + 3-link arm, 3000 SVs, 256 waypoints, Adam — float32 model; the graphed
    step against the autograd step (same arithmetic; float32 Adam amplifies rounding, hence the looser gate)."""
    from diffco_b200 import DiffCo
    from diffco_b200 import kernel as K
    from diffco_b200 import optim as OPT

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
    options = {"n_waypoints": 256, "maxiter": 40, "history": False, "max_move_weight": 10, "collision_weight": 10,
               "joint_limit_weight": 10, "safety_bias": 0.0, "max_speed": 0.3, "optimizer": torch.optim.Adam,
               "optimizer_params": {"lr": 0.02}, "dense_check": False}
    mask = torch.ones(256, dtype=torch.bool)
    mask[[0, -1]] = False
    ref = OPT.Weighted(robot, dc, dict(options)).step(init.clone(), mask=mask)
    fused = OPT.Weighted(robot, dc, dict(options, fused=True)).step(init.clone(), mask=mask)
    err = P.rel_to_max(fused.x, ref.x)
    print(f"cfg-4 256 waypoints x 40 Adam steps: graphed {fused.misc['time']*1e3:.1f} ms, autograd {ref.misc['time']*1e3:.1f} ms, "
          f"max rel diff {err:.2e}")
    assert err <= 2e-3


def test_adam_traj_optimize_graphed_matches_autograd_and_reference(problem):
    """adam_traj_optimize(options['fused'] = True): the reference's bookkeeping (optim.py:86-163) on the CUDA-graph-replayed
    step; same record as the autograd run (float64) and the reference's golden record within its 1e-4 gate."""
    from diffco_b200 import optim as OPT

    g, robot, dc, start, target, init, opts = problem
    base = OPT.adam_traj_optimize(robot, dc.poly_score, start, target, dict(opts, init_solution=init.clone()))
    rec = OPT.adam_traj_optimize(robot, dc.poly_score, start, target, dict(opts, init_solution=init.clone(), fused=True))
    assert rec["success"] == base["success"] and rec["cnt_check"] == base["cnt_check"]
    assert np.abs(np.array(rec["solution"]) - np.array(base["solution"])).max() <= 1e-9
    assert abs(rec["cost"] - base["cost"]) <= 1e-9 * max(1.0, abs(base["cost"]))
    assert np.abs(np.array(rec["solution"]) - g["adam_solution"]).max() <= 1e-4
    with pytest.raises(ValueError):
        OPT.adam_traj_optimize(robot, lambda q: dc.poly_score(q), start, target, dict(opts, init_solution=init.clone(), fused=True))


@pytest.mark.parametrize("mode", ["autograd", "autograd_dense", "graphed", "graphed_dense"])
def test_weighted_step_matches_the_reference_record(mode, cuda_device):
    """Weighted.step on the CUDA path against the waypoints the UNMODIFIED reference's Weighted.step produced
    (tests/golden/weighted.npz: Baxter arm, float64, 25 Adam iterations, with and without dense_check)."""
    from diffco_b200 import DiffCo
    from diffco_b200 import kernel as K
    from diffco_b200 import optim as OPT

    g = np.load(os.path.join(GOLD, "weighted.npz"))
    robot = P.make_robot("baxter")
    dc = DiffCo(kernel_func=K.RQKernel(10.0), transform=robot.fkine, beta=1.0)
    dc.train(T64(g["X"]), T64(g["y"]), max_iteration=len(g["X"]))
    assert dc.support_index.tolist() == g["idx"].tolist()
    dc.fit_poly(K.Polyharmonic(1, 1.0), target="label")
    dense = mode.endswith("_dense")
    cw, mmw, jlw = (float(v) for v in g["weights"])
    options = {"n_waypoints": 12, "maxiter": int(g["maxiter"]), "history": True, "max_move_weight": mmw, "collision_weight": cw,
               "joint_limit_weight": jlw, "safety_bias": float(g["safety_bias"]), "max_speed": float(g["max_speed"]),
               "optimizer": torch.optim.Adam, "optimizer_params": {"lr": float(g["lr"])}, "dense_check": dense,
               "fused": mode.startswith("graphed")}
    res = OPT.Weighted(robot, dc, options).step(T64(g["init"]), mask=torch.from_numpy(g["mask"]))
    assert len(res.misc["path_history"]) == int(g[f"steps_dense{int(dense)}"])
    assert P.rel_to_max(res.x, T64(g[f"x_dense{int(dense)}"])) <= 1e-6
