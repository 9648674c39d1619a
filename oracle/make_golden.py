"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference (imported from
/root/reference through oracle/ref_loader.py) on seeded inputs.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden            # in the build container (needs /root/reference)

The fixtures hold inputs AND the reference's outputs (float64), so the GPU box — which has no /root/reference — can
check both the oracle restatement (tests/test_oracle_vs_golden.py, CPU) and the CUDA path (tests/test_gpu_*.py).
The reference ships no golden vectors of its own for this path, so these are the pin.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def gen_kernels(ns):
    g = torch.Generator().manual_seed(11)
    x = torch.randn(5, 3, 2, generator=g, dtype=torch.float64)
    s = torch.randn(7, 3, 2, generator=g, dtype=torch.float64)
    s[2] = x[1]  # an exact coincidence: r == 0 (Polyharmonic sub-gradient / even-k NaN->0 edge)
    out = {"x": _np(x), "s": _np(s)}
    K = ns.kernel
    cases = {
        "rq_g10_p2": K.RQKernel(10.0), "rq_g3_p3": K.RQKernel(3.0, 3), "rq_g1_p1": K.RQKernel(1.0, 1),
        "ph_k1_e1": K.Polyharmonic(1, 1.0), "ph_k3_e05": K.Polyharmonic(3, 0.5), "ph_k2_e1": K.Polyharmonic(2, 1.0),
        "ph_k1_e001": K.Polyharmonic(1, 0.01),
    }
    for name, k in cases.items():
        out[name] = _np(k(x, s))
        try:  # rank-deficient query: shape quirks (kernel.py:18-27); Polyharmonic's .view raises here (kernel.py:76)
            out[name + "_single"] = _np(k(x[0], s))
        except TypeError:
            pass
    mq = K.MultiQuadratic(0.7)
    out["mq_e07"] = _np(mq(x.reshape(5, -1), s.reshape(7, -1)))
    out["mq_e07_single"] = _np(mq(x.reshape(5, -1)[0], s.reshape(7, -1)))
    # autograd gradients of sum(K w) wrt x for the three inference kernels
    w = torch.randn(7, generator=g, dtype=torch.float64)
    out["w"] = _np(w)
    for name, k in (("rq_g10_p2", cases["rq_g10_p2"]), ("ph_k1_e1", cases["ph_k1_e1"]), ("ph_k3_e05", cases["ph_k3_e05"])):
        xv = x.clone().requires_grad_(True)
        (k(xv, s) @ w).sum().backward()
        out[name + "_gradx"] = _np(xv.grad)
    xv = x.reshape(5, -1).clone().requires_grad_(True)
    (mq(xv, s.reshape(7, -1)) @ w).sum().backward()
    out["mq_e07_gradx"] = _np(xv.grad)
    np.savez_compressed(os.path.join(OUT, "kernels.npz"), **out)


def gen_fk(ns):
    g = torch.Generator().manual_seed(12)
    M = ns.model
    out = {}

    def sample(limits, n=6):
        lim = limits.double()
        return torch.rand(n, lim.shape[0], generator=g, dtype=torch.float64) * (lim[:, 1] - lim[:, 0]) + lim[:, 0]

    def record(name, robot, q, dtype=torch.float64, module="model"):
        qv = q.to(dtype).clone().requires_grad_(True)
        pts = robot.fkine(qv)
        wts = torch.randn(pts.shape, generator=g, dtype=torch.float64).to(dtype)
        (pts * wts).sum().backward()
        out[name + "_q"], out[name + "_x"], out[name + "_gx"], out[name + "_gq"] = _np(q), _np(pts.double()), _np(wts.double()), _np(qv.grad.double())

    for dof, L in ((2, 1.0), (3, [1.0, 0.8, 0.6]), (7, 1.0)):
        r = M.RevolutePlanarRobot(L, 0.3, dof=dof) if not isinstance(L, list) else M.RevolutePlanarRobot(L, 0.3)
        record(f"planar{dof}", r, sample(r.limits))
    parts = [("box", (0.5, 0.2), (1, 1)), ("box", (-0.4, 0.3), (1, 1)), ("box", (0.1, -0.6), (1, 1)),
             ("box", (-0.3, -0.2), (1, 1)), ("box", (0.7, 0.7), (1, 1))]
    r = M.RigidPlanarBody(parts)
    # rot_2d allocates float32 (utils.py:41): the reference's SE(2) map only runs in float32
    record("se2", r, sample(r.limits), dtype=torch.float32)
    # RigidBody.__init__ needs trimesh; fkine only needs dof/keypoints (model.py:156-159)
    r = object.__new__(M.RigidBody)
    r.dof = 6
    c = [[sx * 0.5, sy * 0.3, sz * 0.2] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]
    r.keypoints = torch.FloatTensor(c).T
    r.limits = torch.FloatTensor([[-10, 10]] * 3 + [[-np.pi, np.pi]] * 3)
    record("se3", r, sample(r.limits), dtype=torch.float32)  # float32 keypoints, matmul does not promote (model.py:158)
    r = M.BaxterLeftArmFK()
    record("baxter", r, sample(r.limits))
    r = M.BaxterRightArmFK()
    record("baxter_right", r, sample(r.limits))
    r = M.BaxterDualArmFK()  # float32 only: arm_bases is float32 and torch.matmul does not promote (model.py:376)
    record("baxter_dual", r, sample(r.limits), dtype=torch.float32)
    r = M.PandaFK()
    record("panda", r, sample(r.limits))
    r = ns.robot_fkine.PandaFK()  # older twin: 5 control points (robot_fkine.py:428-444)
    record("panda5", r, sample(r.limits))
    r = M.DualPandaFK()
    record("dual_panda", r, sample(r.limits))
    np.savez_compressed(os.path.join(OUT, "fk.npz"), **out)


def _circle_labels(pts):
    hit = torch.zeros(len(pts), dtype=torch.bool)
    for (cx, cy), rad in (((3.0, 2.0), 2.0), ((-2.0, 3.0), 0.8)):
        hit |= ((pts - torch.tensor([cx, cy], dtype=torch.float64)).norm(dim=2) < rad).any(dim=1)
    return hit.double() * 2 - 1


def gen_perceptron(ns):
    """Train the reference DiffCo on small planar problems; record the selected indices, gains, hypothesis, fit_poly
    nodes, and score / poly_score values + autograd gradients on held-out queries."""
    M, K, P = ns.model, ns.kernel, ns.kernel_perceptrons
    out = {}
    for tag, dof, n_train, seed in (("p2", 2, 400, 21), ("p7", 7, 700, 22)):
        g = torch.Generator().manual_seed(seed)
        robot = M.RevolutePlanarRobot(1.0, 0.3, dof=dof)
        X = (torch.rand(n_train, dof, generator=g, dtype=torch.float64) * 2 - 1) * np.pi
        if dof == 2:
            pts = robot.fkine(X)
            y = (((pts - torch.tensor([1.2, 0.8], dtype=torch.float64)).norm(dim=2) < 0.7).any(1)).double() * 2 - 1
        else:
            y = _circle_labels(robot.fkine(X))
        dc = P.DiffCo(kernel_func=K.RQKernel(10.0), transform=robot.fkine, beta=1.0)
        dc.train(X, y, max_iteration=n_train)
        # indices of the supports inside X (rows are unique)
        idx = torch.tensor([int(torch.where((X == sp).all(1))[0][0]) for sp in dc.support_points])
        dc.fit_poly(K.Polyharmonic(1, 1.0), target="label")
        Q = (torch.rand(24, dof, generator=g, dtype=torch.float64) * 2 - 1) * np.pi
        Q[0] = dc.support_points[3]  # a query that coincides with a support: r == 0 for Polyharmonic
        qv = Q.clone().requires_grad_(True)
        s = dc.score(qv)
        s.sum().backward()
        gs = qv.grad.clone()
        qv = Q.clone().requires_grad_(True)
        ps = dc.poly_score(qv)
        ps.sum().backward()
        gp = qv.grad.clone()
        out.update({f"{tag}_X": _np(X), f"{tag}_y": _np(y), f"{tag}_idx": _np(idx), f"{tag}_gains": _np(dc.gains),
                    f"{tag}_hyp": _np(dc.hypothesis), f"{tag}_K": _np(dc.kernel_matrix), f"{tag}_nodes": _np(dc.rbf_nodes),
                    f"{tag}_Q": _np(Q), f"{tag}_score": _np(s), f"{tag}_score_grad": _np(gs), f"{tag}_poly": _np(ps),
                    f"{tag}_poly_grad": _np(gp), f"{tag}_single_score": _np(dc.score(Q[5])),
                    f"{tag}_single_poly": _np(dc.poly_score(Q[5]))})
        if tag == "p7":
            # active-learning round: jump-start state for (old supports + novel points) (kernel_perceptrons.py:222-269)
            novel = (torch.rand(150, dof, generator=g, dtype=torch.float64) * 2 - 1) * np.pi
            Xu = torch.cat([dc.support_points, novel])
            yu = torch.cat([dc.y, _circle_labels(robot.fkine(novel))])
            exist = torch.zeros(len(Xu), dtype=torch.bool)
            exist[: len(dc.support_points)] = True
            gains0, _, _, K0, h0, _ = dc.jump_start_initialize(Xu, yu, exist)
            dc.train(Xu, yu, update=True, exist_mask=exist, max_iteration=len(Xu))
            idx2 = torch.tensor([int(torch.where((Xu == sp).all(1))[0][0]) for sp in dc.support_points])
            out.update({"p7u_X": _np(Xu), "p7u_y": _np(yu), "p7u_exist": _np(exist), "p7u_gains0": _np(gains0), "p7u_h0": _np(h0),
                        "p7u_K0": _np(K0), "p7u_idx": _np(idx2), "p7u_gains": _np(dc.gains), "p7u_hyp": _np(dc.hypothesis)})
    np.savez_compressed(os.path.join(OUT, "perceptron.npz"), **out)


def gen_multiclass(ns):
    """Legacy 4-class MultiDiffCo on Baxter (deprecated/MultiDiffCo.py) with the FKKernel shim."""
    M, K = ns.model, ns.kernel
    g = torch.Generator().manual_seed(31)
    robot = M.BaxterLeftArmFK()
    lim = robot.limits.double()
    X = torch.rand(160, 7, generator=g, dtype=torch.float64) * (lim[:, 1] - lim[:, 0]) + lim[:, 0]
    pts = robot.fkine(X)
    centers = torch.tensor([[0.6, 0.2, 0.3], [0.3, -0.5, 0.6], [0.8, 0.0, -0.1], [0.2, 0.6, 0.1]], dtype=torch.float64)
    radii = torch.tensor([0.45, 0.4, 0.4, 0.35], dtype=torch.float64)
    Y = torch.stack([((pts - centers[c]).norm(dim=2) < radii[c]).any(1).double() * 2 - 1 for c in range(4)], dim=1)
    fkk = ns.FKKernelShim(robot.fkine, K.RQKernel(10.0))
    mdc = ns.legacy_MultiDiffCo.MultiDiffCo(None, kernel_func=fkk, beta=1.0)
    mdc.train(X, Y, max_iteration=len(X))
    idx = torch.tensor([int(torch.where((X == sp).all(1))[0][0]) for sp in mdc.support_points])
    # default rbf kernel MultiQuadratic(1) (deprecated/MultiDiffCo.py:138); Polyharmonic(1) has a zero diagonal and the
    # block-zeroing of :143-150 makes the system singular on this data
    mdc.fit_poly(kernel_func=None, target="label", fkine=robot.fkine)
    Q = torch.rand(20, 7, generator=g, dtype=torch.float64) * (lim[:, 1] - lim[:, 0]) + lim[:, 0]
    go = torch.randn(20, 4, generator=g, dtype=torch.float64)
    qv = Q.clone().requires_grad_(True)
    s = mdc.score(qv)
    (s * go).sum().backward()
    gs = qv.grad.clone()
    qv = Q.clone().requires_grad_(True)
    r = mdc.rbf_score(qv)
    (r * go).sum().backward()
    gr = qv.grad.clone()
    jac = torch.autograd.functional.jacobian(lambda q: mdc.rbf_score(q).sum(0), Q)  # (C, B, D)
    np.savez_compressed(os.path.join(OUT, "multiclass.npz"), X=_np(X), Y=_np(Y), idx=_np(idx), gains=_np(mdc.gains),
                        hyp=_np(mdc.hypothesis), nodes=_np(mdc.rbf_nodes), Q=_np(Q), go=_np(go), score=_np(s),
                        score_grad=_np(gs), rbf=_np(r), rbf_grad=_np(gr), rbf_jac=_np(jac.permute(1, 0, 2)))


def gen_optim_replay(ns):
    """Run the reference optimisers (diffco/optim.py) with the reference perceptron as ``dist_est`` and record every
    query it receives with the value and the gradient autograd produced.  The GPU tests replay the queries through the
    CUDA ``dist_est`` (the optimisers themselves live in /root/reference and do not travel)."""
    M, K, P, OPT = ns.model, ns.kernel, ns.kernel_perceptrons, ns.optim
    g = torch.Generator().manual_seed(41)
    robot = M.RevolutePlanarRobot(1.0, 0.3, dof=7)
    X = (torch.rand(900, 7, generator=g, dtype=torch.float64) * 2 - 1) * np.pi
    y = _circle_labels(robot.fkine(X))
    dc = P.DiffCo(kernel_func=K.RQKernel(10.0), transform=robot.fkine, beta=1.0)
    dc.train(X, y, max_iteration=len(X))
    dc.fit_poly(K.Polyharmonic(1, 1.0), target="label")
    idx = torch.tensor([int(torch.where((X == sp).all(1))[0][0]) for sp in dc.support_points])
    calls = []

    def dist_est(p):
        out = dc.poly_score(p)
        calls.append((p.detach().clone(), out.detach().clone()))
        return out

    start = torch.tensor([-2.0, -0.4, 0.3, -0.2, 0.1, 0.2, -0.1], dtype=torch.float64)
    target = torch.tensor([1.6, 0.5, -0.3, 0.4, -0.2, 0.1, 0.3], dtype=torch.float64)
    init = torch.from_numpy(np.linspace(start.numpy(), target.numpy(), 12))
    opts = {"N_WAYPOINTS": 12, "NUM_RE_TRIALS": 1, "MAXITER": 15, "safety_margin": -0.3, "max_speed": 0.6, "seed": 1234,
            "history": False, "extra_optimizer_options": {"lr": 0.05}, "init_solution": init.clone()}
    rec_adam = OPT.adam_traj_optimize(robot, dist_est, start, target, dict(opts))
    n_adam = len(calls)
    opts2 = dict(opts)
    opts2["extra_optimizer_options"] = {"ftol": 1e-4, "disp": False}
    opts2["MAXITER"] = 6
    opts2["init_solution"] = init.clone()
    rec_slsqp = OPT.givengrad_traj_optimize(robot, dist_est, start, target, opts2)
    # keep a bounded subset of the queries, with autograd gradients of sum(clamp(score - margin, min=0))
    keep = list(range(0, n_adam, 3))[:6] + list(range(n_adam, len(calls), max(1, (len(calls) - n_adam) // 6)))[:6]
    out = {"X": _np(X), "y": _np(y), "idx": _np(idx), "gains": _np(dc.gains), "nodes": _np(dc.rbf_nodes),
           "support_points": _np(dc.support_points), "n_calls": np.array([n_adam, len(calls) - n_adam]),
           "adam_solution": np.array(rec_adam["solution"]), "adam_cost": np.array(rec_adam["cost"]),
           "slsqp_solution": np.array(rec_slsqp["solution"]), "slsqp_cost": np.array(rec_slsqp["cost"])}
    for j, ci in enumerate(keep):
        p, val = calls[ci]
        pv = p.clone().requires_grad_(True)
        sc = dc.poly_score(pv)
        torch.clamp(sc - opts["safety_margin"], min=0).sum().backward()
        out[f"call{j}_p"], out[f"call{j}_score"], out[f"call{j}_grad"] = _np(p), _np(val), _np(pv.grad)
    out["n_kept"] = np.array(len(keep))
    # the SLSQP constraint Jacobian path: jacobian(vectorize=True) of per-segment sums (optim.py:190-218)
    p = torch.tensor(rec_adam["solution"], dtype=torch.float64)

    def con(pp):
        dense = ns.utils.dense_path(pp, opts["max_speed"])
        cost = -(dc.poly_score(dense[1:-1]) - opts["safety_margin"])
        cost = torch.clamp_(cost, max=0).reshape(-1)
        n_seg, n_pt = len(pp) - 1, len(dense) - 2
        mult = n_pt // n_seg + (1 if n_pt % n_seg else 0)
        if n_seg * mult - n_pt:
            cost = torch.cat([cost, torch.zeros(n_seg * mult - n_pt, dtype=cost.dtype)])
        return cost.reshape(n_seg, -1).sum(dim=1)

    jac = torch.autograd.functional.jacobian(con, p.clone().requires_grad_(True), create_graph=False, strict=False,
                                             vectorize=True, strategy="reverse-mode")
    out["con_p"], out["con_val"], out["con_jac"] = _np(p), _np(con(p)), _np(jac)
    out["safety_margin"], out["max_speed"] = np.array(opts["safety_margin"]), np.array(opts["max_speed"])
    np.savez_compressed(os.path.join(OUT, "optim_replay.npz"), **out)


def gen_weighted(ns):
    """The reference's ``Weighted.step`` (optim.py:686-761) on a Baxter arm (the only FK classes whose ``fkine`` takes the
    ``reuse=`` argument that method passes), with and without ``dense_check``.  The current ``DiffCo`` has no
    ``rbf_score``, so the harness binds the alias the method looks up.  On the CPU the method's
    ``path_history`` entries all alias the final tensor (``p.cpu()`` is ``p``), so only the final path and the number of
    iterations are recorded."""
    M, K, P, OPT = ns.model, ns.kernel, ns.kernel_perceptrons, ns.optim
    g = torch.Generator().manual_seed(43)
    robot = M.BaxterLeftArmFK()
    lim = robot.limits.double()
    X = torch.rand(900, 7, generator=g, dtype=torch.float64) * (lim[:, 1] - lim[:, 0]) + lim[:, 0]
    cp = robot.fkine(X)
    y = ((cp - torch.tensor([0.6, 0.4, 0.3], dtype=cp.dtype)).norm(dim=2) < 0.35).any(1).double() * 2 - 1
    dc = P.DiffCo(kernel_func=K.RQKernel(10.0), transform=robot.fkine, beta=1.0)
    dc.train(X, y, max_iteration=len(X))
    dc.fit_poly(K.Polyharmonic(1, 1.0), target="label")
    dc.rbf_score = dc.poly_score
    idx = torch.tensor([int(torch.where((X == sp).all(1))[0][0]) for sp in dc.support_points])
    init = torch.from_numpy(np.linspace(X[0].numpy(), X[1].numpy(), 12))
    mask = torch.ones(12, dtype=torch.bool)
    mask[[0, -1]] = False
    out = {"X": _np(X), "y": _np(y), "idx": _np(idx), "gains": _np(dc.gains), "nodes": _np(dc.rbf_nodes),
           "support_points": _np(dc.support_points), "init": _np(init), "mask": _np(mask),
           "maxiter": np.array(25), "weights": np.array([10.0, 10.0, 10.0]), "safety_bias": np.array(0.3),
           "max_speed": np.array(0.3), "lr": np.array(0.05)}
    for dense in (False, True):
        options = {"n_waypoints": 12, "maxiter": 25, "history": True, "max_move_weight": 10, "collision_weight": 10,
                   "joint_limit_weight": 10, "safety_bias": 0.3, "max_speed": 0.3, "optimizer": torch.optim.Adam,
                   "optimizer_params": {"lr": 0.05}, "dense_check": dense}
        res = OPT.Weighted(robot, dc, options).step(init.clone(), mask=mask)
        out[f"x_dense{int(dense)}"] = _np(res.x)
        out[f"steps_dense{int(dense)}"] = np.array(len(res.misc["path_history"]))
    np.savez_compressed(os.path.join(OUT, "weighted.npz"), **out)


def gen_checkers(ns):
    """The reference's high-level checker (diffco/collision_checkers.py: ForwardKinematicsDiffCo.fit / update-style refit /
    verify / collision_score) on a 7-link planar arm with a synthetic ground truth (circle obstacles).  The module's
    geometry back ends (yourdfpy / trimesh / python-fcl / cuRobo / ROS) are stubbed: with an injected gt_check_func and a
    robot object that provides the members the checker touches, none of them is reached."""
    import importlib
    import sys
    import types

    M = ns.model
    ci = types.ModuleType("diffco.collision_interfaces")
    for name in ("RobotInterfaceBase", "URDFRobot", "MultiURDFRobot", "ROSRobotEnv", "CuRoboRobot", "CuRoboCollisionWorldEnv",
                 "ShapeEnv", "PCDEnv"):
        setattr(ci, name, type(name, (), {}))
    ci.robot_description_folder = ""
    sys.modules["diffco.collision_interfaces"] = ci
    pkg = sys.modules["diffco"]
    pkg.collision_interfaces, pkg.model, pkg.kernel = ci, ns.model, ns.kernel
    cc = importlib.import_module("diffco.collision_checkers")

    planar = M.RevolutePlanarRobot(1.0, 0.3, dof=7)

    class Link:
        def __init__(self, name):
            self.name = name

        def joint_trans(self):
            return torch.ones(3)

    class FakeRobot(ci.RobotInterfaceBase):
        """What ForwardKinematicsDiffCo touches of a URDFRobot: _bodies, _n_dofs, joint_limits, rand_configs,
        compute_forward_kinematics_all_links (link name -> [(position (B, 3), rotation)])."""

        def __init__(self):
            self._bodies = [Link(f"l{i}") for i in range(7)]
            self._n_dofs = 7
            self.joint_limits = planar.limits.double()
            self.gen = torch.Generator().manual_seed(99)

        def rand_configs(self, n):
            lo, hi = self.joint_limits[:, 0], self.joint_limits[:, 1]
            return torch.rand(n, 7, generator=self.gen, dtype=torch.float64) * (hi - lo) + lo

        def compute_forward_kinematics_all_links(self, q, return_collision=False):
            pts = planar.fkine(q)  # (B, 7, 2)
            pos = torch.cat([pts, torch.zeros_like(pts[..., :1])], dim=-1)
            return {f"l{i}": [(pos[:, i], None)] for i in range(7)}

        def collision(self, q, other=None):
            raise AssertionError("the ground truth is injected")

    def gt(q):
        return (_circle_labels(planar.fkine(q)) > 0).to(q.dtype)

    g = torch.Generator().manual_seed(31)
    lim = planar.limits.double()
    X = torch.rand(900, 7, generator=g, dtype=torch.float64) * (lim[:, 1] - lim[:, 0]) + lim[:, 0]
    chk = cc.ForwardKinematicsDiffCo(robot=FakeRobot(), gt_check_func=gt, device="cpu")
    torch.manual_seed(5)
    acc = chk.fit(q=X.clone(), verify_ratio=0.2)
    dc = chk.perceptron
    out = {"X": _np(X), "limits": _np(lim), "fit_rates": np.array([float(v) for v in acc]), "fit_bias": _np(chk.safety_bias),
           "fit_support_points": _np(dc.support_points), "fit_nodes": _np(dc.rbf_nodes), "fit_q_verify": _np(chk.q_verify)}
    Q = torch.rand(64, 7, generator=g, dtype=torch.float64) * (lim[:, 1] - lim[:, 0]) + lim[:, 0]
    out["Q"] = _np(Q)
    out["fit_collision_score"] = _np(chk.collision_score(Q.reshape(4, 16, 7)))
    out["fit_collision_score_links"] = _np(chk.collision_score(q_link_pos=chk.tensorized_fkine(Q)))  # (B, 3, L): the layout of support_transformed
    # active-learning refit (update() builds q and exist_mask from the RNG; the same refit with explicit inputs):
    novel = torch.rand(200, 7, generator=g, dtype=torch.float64) * (lim[:, 1] - lim[:, 0]) + lim[:, 0]
    Xu = torch.cat([novel, dc.support_points], dim=0)
    exist = torch.zeros(len(Xu), dtype=torch.bool)
    exist[-len(dc.support_points):] = True
    chk.fit(Xu.clone(), update=True, exist_mask=exist, verify_ratio=0)  # q_verify = robot.rand_configs(100)
    out.update({"upd_X": _np(Xu), "upd_exist": _np(exist), "upd_support_points": _np(dc.support_points),
                "upd_nodes": _np(dc.rbf_nodes), "upd_bias": _np(chk.safety_bias),
                "upd_q_verify": _np(FakeRobot().rand_configs(100)),  # the first draw of a fresh generator: what fit() consumed
                "upd_collision_score": _np(chk.collision_score(Q))})
    ver = chk.verify(q_verify=Q)
    out["upd_verify_rates"] = np.array([float(v) for v in ver])
    np.savez_compressed(os.path.join(OUT, "checkers.npz"), **out)


def gen_line_temporal(ns):
    """TemporalFKKernel / LineFKKernel / LineKernel (kernel.py:145-202) on the 7-link planar arm: kernel matrices and the
    autograd gradient of sum_n w_n k(x, s_n) w.r.t. the raw rows ([q | t] resp. [q_a | q_b])."""
    K, M = ns.kernel, ns.model
    g = torch.Generator().manual_seed(41)
    robot = M.RevolutePlanarRobot(1.0, 0.3, 7)
    fk = lambda q: robot.fkine(q)
    out = {}
    # temporal: rows [q (7) | t]
    x = torch.cat([(torch.rand(9, 7, generator=g, dtype=torch.float64) * 2 - 1) * np.pi, torch.rand(9, 1, generator=g, dtype=torch.float64)], 1)
    s = torch.cat([(torch.rand(33, 7, generator=g, dtype=torch.float64) * 2 - 1) * np.pi, torch.rand(33, 1, generator=g, dtype=torch.float64)], 1)
    s[4] = x[2]
    s[5, :7] = x[3, :7]  # same configuration, different time
    w = torch.randn(33, generator=g, dtype=torch.float64)
    out.update(t_x=_np(x), t_s=_np(s), w=_np(w))
    for name, (gx, px, gt, pt, al) in {"t_a": (10.0, 2, 5.0, 2, 0.5), "t_b": (3.0, 1, 20.0, 3, 2.0)}.items():
        k = K.TemporalFKKernel(fk, K.RQKernel(gx, px), K.RQKernel(gt, pt), alpha=al)
        xv = x.clone().requires_grad_(True)
        km = k(xv, s)
        (km @ w).sum().backward()
        out[name + "_params"] = np.array([gx, px, gt, pt, al])
        out[name + "_K"] = _np(km)
        out[name + "_grad"] = _np(xv.grad)
        out[name + "_K_single"] = _np(k(x[0], s))
    # line: rows [q_a | q_b]
    xl = (torch.rand(9, 14, generator=g, dtype=torch.float64) * 2 - 1) * np.pi
    sl = (torch.rand(33, 14, generator=g, dtype=torch.float64) * 2 - 1) * np.pi
    sl[7] = xl[1]
    out.update(l_x=_np(xl), l_s=_np(sl))
    k = K.LineFKKernel(fk, K.RQKernel(10.0))
    xv = xl.clone().requires_grad_(True)
    km = k(xv, sl)
    (km @ w).sum().backward()
    out.update(l_K=_np(km), l_grad=_np(xv.grad), l_K_single=_np(k(xl[0], sl)))
    lk = K.LineKernel(K.RQKernel(2.0))
    out["lk_K"] = _np(lk(xl, sl))
    np.savez_compressed(os.path.join(OUT, "line_temporal.npz"), **out)


def _reference_urdf_tree(path, base=None):
    """The reference's own RigidBody tree (rigid_body.py, imported unmodified) built the way URDFRobot.__init__ builds it
    (urdf_interface.py:370-403,556-600) — except that the file is read with xml.etree here because yourdfpy is not
    installed.  Returns (bodies, fk) with fk(q) -> {link: (translation, rotation)} as
    compute_forward_kinematics_all_links does (urdf_interface.py:517-553)."""
    import importlib
    import types
    import xml.etree.ElementTree as ET

    if "diffco.collision_interfaces" not in sys.modules or not hasattr(sys.modules["diffco.collision_interfaces"], "__path__"):
        pkg = types.ModuleType("diffco.collision_interfaces")
        pkg.__path__ = [os.path.join(ref_loader.REFERENCE_ROOT, "diffco", "collision_interfaces")]
        sys.modules["diffco.collision_interfaces"] = pkg
    RB = importlib.import_module("diffco.collision_interfaces.rigid_body")
    SVA = importlib.import_module("diffco.collision_interfaces.spatial_vector_algebra")
    root = ET.parse(path).getroot()
    f3 = lambda el, key: [float(v) for v in el.get(key, "0 0 0").split()]
    jmap = {j.find("child").get("link"): j for j in root.findall("joint")}
    jname = {j.get("name"): j for j in root.findall("joint")}
    bodies, controlled, mimics = [], [], {}
    for idx, link in enumerate(root.findall("link")):
        j = jmap.get(link.get("name"))
        prm = {"link_idx": idx, "link_name": link.get("name"), "joint_limits": None, "joint_mimic": None}
        if j is None:
            prm.update(joint_rot_angles=torch.zeros(3), joint_trans=torch.zeros(3), joint_name="base_joint", joint_type="fixed",
                       joint_axis=torch.zeros((1, 3)))
        else:
            o = j.find("origin")
            prm.update(joint_rot_angles=torch.tensor(f3(o, "rpy") if o is not None else [0.0] * 3, dtype=torch.float32),
                       joint_trans=torch.tensor(f3(o, "xyz") if o is not None else [0.0] * 3, dtype=torch.float32),
                       joint_name=j.get("name"), joint_type=j.get("type"), joint_axis=torch.zeros((1, 3)))
            if j.get("type") != "fixed":
                ax = j.find("axis")
                prm["joint_axis"] = torch.tensor(f3(ax, "xyz") if ax is not None else [1.0, 0, 0], dtype=torch.float32).reshape(1, 3)
                m = j.find("mimic")
                if m is not None:
                    prm["joint_mimic"] = types.SimpleNamespace(joint=m.get("joint"), multiplier=float(m.get("multiplier", 1.0)),
                                                               offset=float(m.get("offset", 0.0)))
        b = RB.RigidBody(rigid_body_params=prm)
        if b.joint_type != "fixed":
            if b.joint_mimic is None:
                controlled.append(idx)
            else:
                mimics.setdefault(jname[b.joint_mimic.joint].find("child").get("link"), []).append(b.name)
        bodies.append(b)
    by_name = {b.name: b for b in bodies}
    for b in bodies:
        if b.joint_name != "base_joint":
            parent = by_name[jname[b.joint_name].find("parent").get("link")]
            b.set_parent(parent)
            parent.add_child(b)
    base = torch.eye(4) if base is None else base
    base_tf = SVA.CoordinateTransform(base[:3, :3], base[:3, 3])

    def fk(q):
        q_dict = {}
        for i, bi in enumerate(controlled):
            q_dict[bodies[bi].name] = q[:, i].unsqueeze(1)
            for mn in mimics.get(bodies[bi].name, []):
                q_dict[mn] = q_dict[bodies[bi].name]
        poses = bodies[0].forward_kinematics(q_dict, False)
        return {k: base_tf.multiply_transform(v[0]) for k, v in poses.items()}

    return bodies, controlled, fk


def gen_urdf(ns):
    """URDF-tree forward kinematics (rigid_body.py:86-141 via urdf_interface.py:517-553) on the synthetic test URDFs and,
    for the reference's own Panda description, on the joint program this package compiled from it (the file itself does
    not travel: the golden holds the compiled dc_fk_desc bytes, q and the reference's outputs).  float32, as the reference
    computes (its rotation constructors allocate float32)."""
    import ctypes

    from diffco_b200.collision_interfaces import URDFRobot  # descriptor bytes + feature-link order only

    out = {}
    cases = {"arm7": (os.path.join(ROOT, "tests", "data", "arm7_gripper.urdf"), None),
             "torso": (os.path.join(ROOT, "tests", "data", "torso_two_arms.urdf"),
                       torch.tensor([[0.0, -1.0, 0.0, 0.3], [1.0, 0.0, 0.0, -0.2], [0.0, 0.0, 1.0, 0.1], [0, 0, 0, 1.0]])),
             "panda": (os.path.join(ref_loader.REFERENCE_ROOT, "diffco", "robot_data", "panda_description", "urdf", "panda.urdf"), None)}
    g = torch.Generator().manual_seed(77)
    for name, (path, base) in cases.items():
        bodies, controlled, fk = _reference_urdf_tree(path, base)
        lim = torch.zeros(len(controlled), 2)
        for i, bi in enumerate(controlled):
            b = bodies[bi]
            lim[i] = torch.tensor([-np.pi, np.pi]) if b.joint_type != "continuous" else torch.tensor([-2 * np.pi, 2 * np.pi])
        q = (torch.rand(16, len(controlled), generator=g) * (lim[:, 1] - lim[:, 0]) + lim[:, 0]) * 0.6
        q[0] = 0
        unique = [b.name for b in bodies if torch.any(b.joint_trans() != 0)]  # collision_checkers.py:356-358
        qv = q.clone().requires_grad_(True)
        poses = fk(qv)
        feat = torch.stack([poses[n].translation().expand(len(q), 3) for n in unique], dim=-1)  # (B, 3, L): tensorized_fkine_single_robot
        gx = torch.randn(feat.shape, generator=g)
        (feat * gx).sum().backward()
        out[name + "_q"], out[name + "_x"], out[name + "_gx"], out[name + "_gq"] = _np(q), _np(feat), _np(gx), _np(qv.grad)
        out[name + "_unique"] = np.array(unique)
        out[name + "_links"] = np.array(list(poses.keys()))
        # bodies above the first joint come back with batch size 1 (rigid_body.py:126): expand for stacking
        out[name + "_trans"] = _np(torch.stack([poses[k].translation().expand(len(q), 3) for k in poses], 1))
        out[name + "_rot"] = _np(torch.stack([poses[k].rotation().expand(len(q), 3, 3) for k in poses], 1))
        robot = URDFRobot(path, base_transform=base)
        assert robot.unique_position_link_names == unique, (robot.unique_position_link_names, unique)
        out[name + "_desc"] = np.frombuffer(bytes(ctypes.string_at(ctypes.addressof(robot.fk_desc), ctypes.sizeof(robot.fk_desc))),
                                            dtype=np.uint8).copy()
        out[name + "_nodes"] = np.array(robot.node_names)
        out[name + "_limits"] = _np(robot.joint_limits)
    np.savez_compressed(os.path.join(OUT, "urdf.npz"), **out)


def main():
    warnings.filterwarnings("ignore")
    torch.set_num_threads(4)
    os.makedirs(OUT, exist_ok=True)
    ns = ref_loader.load_legacy()
    only = sys.argv[1:]
    for fn in (gen_kernels, gen_fk, gen_perceptron, gen_multiclass, gen_optim_replay, gen_weighted, gen_checkers, gen_line_temporal, gen_urdf):
        if only and fn.__name__ not in only:
            continue
        fn(ns)
        print("wrote", fn.__name__)


if __name__ == "__main__":
    main()
