"""Trajectory optimisers — host-side mirror of the callers of the hot path in the reference's ``diffco/optim.py``.

Same call signatures, option keys and result records as the reference (``adam_traj_optimize`` optim.py:13-163,
``givengrad_traj_optimize`` optim.py:166-321, ``Weighted.step`` optim.py:655-761), so scripts switch by changing the
import.  ``dist_est`` is any callable with the protocol of SURVEY.md §8b — normally ``checker.poly_score`` /
``checker.rbf_score`` / ``checker.score`` of a diffco_b200 perceptron, i.e. the fused CUDA kernel with its analytic
Jacobian behind autograd.  The optimisation drivers themselves are host logic (Adam / scipy SLSQP over ~20 waypoints);
what they spend their time in — dist_est and robot.fkine with gradients — runs on the GPU.

``Weighted.step`` and ``adam_traj_optimize`` additionally have a device-resident fast path (``options['fused'] = True``,
diffco_b200/trajopt.py): the penalty (collision hinge + max-move + joint-limit + path length), its analytic gradient,
the Adam update and ``robot.wrap`` are ONE launch (``dc_traj_step``) after ``dc_score_grad``, both captured in a CUDA graph
that is replayed per iteration.

``trustconstr_traj_optimize`` needs the Hessian of the collision constraint (optim.py:380-391); the fused score gives
first derivatives analytically, so the Hessian is a central difference of that analytic gradient with ALL perturbed
paths evaluated in one ``dist_est`` launch (``_SlsqpProblem.hess_con_collision``).

Not provided: ``gradient_free_traj_optimize`` — raises NotImplementedError rather than silently doing something else.
"""
from __future__ import annotations

import time
from collections import namedtuple

import numpy as np
import torch
from scipy.optimize import NonlinearConstraint, minimize

from . import utils

OptimizerResult = namedtuple("OptimizerResult", ["x", "misc"])

# weights of optim.py:19-22
_DIF_WEIGHT, _MAX_MOVE_WEIGHT, _COLLISION_WEIGHT, _JOINT_LIMIT_WEIGHT = 1, 10, 10, 10


def _cfg_list(t):
    return (t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)).tolist()


def _path_length(robot, p):
    """sum_i |X_{i+1} - X_i|^2 over consecutive control-point sets (optim.py:46-47,238-244)."""
    cp = robot.fkine(p)
    return (cp[1:] - cp[:-1]).square().sum()


def _initial_path(robot, start_cfg, target_cfg, n_waypoints, trial, options):
    """Trial 0: options['init_solution'] or the straight line; later trials: uniform samples inside the joint limits
    (optim.py:56-80).  Returns (path, trivial) — trivial when the given solution has only the two end points."""
    if trial == 0:
        if "init_solution" in options:
            init = options["init_solution"]
            assert isinstance(init, torch.Tensor) and len(init) >= 2
            path = init.clone()
            if len(path) == 2:
                return path, True
        else:
            s = start_cfg.detach().cpu().double().numpy()
            t = target_cfg.detach().cpu().double().numpy()
            path = torch.from_numpy(np.linspace(s, t, num=n_waypoints)).double()
    else:
        lim = robot.limits
        path = torch.rand((n_waypoints, robot.dof)).double() * (lim[:, 1] - lim[:, 0]) + lim[:, 0]
    path[0] = start_cfg
    path[-1] = target_cfg
    return path, False


def _trivial_record(robot, path, start_cfg, target_cfg, seed, t0):
    with torch.no_grad():
        c = _path_length(robot, path[1:-1]).item() if len(path) > 2 else 0.0
    return {"start_cfg": _cfg_list(start_cfg), "target_cfg": _cfg_list(target_cfg), "cnt_check": 0, "cost": c,
            "time": time.time() - t0, "success": True, "seed": seed, "solution": _cfg_list(path)}


def adam_traj_optimize(robot, dist_est, start_cfg, target_cfg, options):
    """optim.py:13-163.  Penalty method with Adam: minimise path length + 10*(collision hinge + max-move + joint-limit)
    over the interior waypoints; restart from random paths up to NUM_RE_TRIALS times; return the best feasible
    (constraint loss <= 1e-2) or else the lowest-loss path, as the reference's record dict."""
    n_waypoints, n_trials, maxiter = options["N_WAYPOINTS"], options["NUM_RE_TRIALS"], options["MAXITER"]
    keep_history = options["history"]
    safety_margin, max_speed = options["safety_margin"], options["max_speed"]
    lr = options.get("extra_optimizer_options", {}).get("lr", 5e-1)
    seed = options["seed"]
    torch.manual_seed(seed)

    lowest = {"loss": np.inf, "obj": np.inf, "p": None, "step": None, "trial": None}
    valid = {"obj": np.inf, "p": None, "step": None, "trial": None}
    histories, cnt_check, found = [], 0, False
    t0 = time.time()
    stepper = None
    if options.get("fused", False):
        # device-resident fast path (diffco_b200/trajopt.py): every iteration is one CUDA-graph replay of
        # dc_score_grad + dc_traj_step; the bookkeeping below is the reference's, fed from the kernel's six terms
        from . import trajopt

        target = trajopt.scorer_of(dist_est)
        if target is None or isinstance(safety_margin, torch.Tensor):
            raise ValueError("options['fused'] = True needs dist_est = <diffco_b200 perceptron>.score / .poly_score / .rbf_score "
                             "and a scalar safety_margin")
    for trial in range(n_trials):
        path, trivial = _initial_path(robot, start_cfg, target_cfg, n_waypoints, trial, options)
        if trivial:
            return _trivial_record(robot, path, start_cfg, target_cfg, seed, t0)
        if options.get("fused", False):
            checker, weights = target
            if stepper is None or stepper.p.shape != path.shape:  # init_solution may have a different length than N_WAYPOINTS
                mask = torch.ones(len(path), dtype=torch.bool)
                mask[[0, -1]] = False  # the end points are fixed
                stepper = trajopt.GraphedPenaltyStep(robot, checker, weights, path.to(checker.device), mask,
                                                     dif_weight=_DIF_WEIGHT, max_move_weight=_MAX_MOVE_WEIGHT,
                                                     collision_weight=_COLLISION_WEIGHT, joint_limit_weight=_JOINT_LIMIT_WEIGHT,
                                                     safety_bias=-float(safety_margin), max_speed=max_speed, lr=lr, wrap=False)
            else:
                stepper.reset(path)
            history = []
            for step in range(maxiter):
                t = stepper.replay()  # [path length, collision, joint limit, max move, constraint, |grad|^2] before the update
                cnt_check += len(path)
                if keep_history:
                    history.append(stepper.p.detach().clone())
                ov, cv = _DIF_WEIGHT * t[0], t[4]
                lv = ov + cv
                if lv < lowest["loss"]:
                    lowest.update(loss=lv, obj=ov, p=stepper.p.detach().clone(), step=step, trial=trial)
                if cv <= 1e-2 and ov < valid["obj"]:
                    valid.update(obj=ov, p=stepper.p.detach().clone(), step=step, trial=trial)
                if cv <= 1e-2 and t[5] ** 0.5 < 1e-4:
                    break
            histories.append(history)
            if valid["p"] is not None:
                found = True
                break
            continue
        p = path.requires_grad_(True)
        opt = torch.optim.Adam([p], lr=lr)
        history = []
        for step in range(maxiter):
            opt.zero_grad()
            collision = torch.clamp(dist_est(p) - safety_margin, min=0).sum()
            cnt_check += len(p)
            cp = robot.fkine(p)
            seg = (cp[1:] - cp[:-1]).square()
            max_move = torch.clamp(seg.sum(dim=2) - max_speed**2, min=0).sum()
            joint_limit = (torch.clamp(robot.limits[:, 0] - p, min=0) + torch.clamp(p - robot.limits[:, 1], min=0)).sum()
            constraint = _COLLISION_WEIGHT * collision + _MAX_MOVE_WEIGHT * max_move + _JOINT_LIMIT_WEIGHT * joint_limit
            objective = _DIF_WEIGHT * seg.sum()
            loss = objective + constraint
            loss.backward()
            p.grad[[0, -1]] = 0.0  # the end points are fixed
            opt.step()
            if keep_history:
                history.append(p.data.clone())
            lv, ov, cv = loss.item(), objective.item(), constraint.item()
            if lv < lowest["loss"]:
                lowest.update(loss=lv, obj=ov, p=p.data.clone(), step=step, trial=trial)
            if cv <= 1e-2 and ov < valid["obj"]:
                valid.update(obj=ov, p=p.data.clone(), step=step, trial=trial)
            if cv <= 1e-2 and torch.norm(p.grad) < 1e-4:
                break
        histories.append(history)
        if valid["p"] is not None:
            found = True
            break
    pick = valid if found else lowest
    return {"start_cfg": _cfg_list(start_cfg), "target_cfg": _cfg_list(target_cfg), "cnt_check": cnt_check,
            "cost": float(pick["obj"]), "time": time.time() - t0, "success": found, "seed": seed,
            "solution": _cfg_list(pick["p"])}


class _SlsqpProblem:
    """Objective and constraints of optim.py:182-255 for one initial path.  Decision variables: the interior waypoints,
    flattened; the end points ride along as constants.  Each callback re-builds the full path, evaluates with autograd
    and caches what its gradient callback needs, exactly one evaluation per scipy call."""

    def __init__(self, robot, dist_est, init_path, safety_margin, max_speed):
        self.robot, self.dist_est = robot, dist_est
        self.first, self.last = init_path[:1].detach().clone(), init_path[-1:].detach().clone()
        self.safety_margin, self.max_speed = safety_margin, max_speed
        self.cnt_check = 0
        self._cost_cache = None
        self._limit_cache = None

    def full_path(self, x):
        mid = torch.as_tensor(np.asarray(x), dtype=torch.float64).reshape(-1, self.robot.dof)
        return torch.cat([self.first, mid, self.last], dim=0).requires_grad_(True)

    # objective --------------------------------------------------------------------------------------------------
    def cost(self, x):
        p = self.full_path(x)
        obj = _path_length(self.robot, p)
        self._cost_cache = (np.array(x, copy=True), p, obj)
        return obj.item()

    def grad_cost(self, x):
        if self._cost_cache is None or not np.allclose(x, self._cost_cache[0]):
            self.cost(x)
        _, p, obj = self._cost_cache
        p.grad = None
        obj.backward(retain_graph=True)
        if p.grad is None:
            return np.zeros(len(x), dtype=np.asarray(x).dtype)
        return p.grad[1:-1].numpy().reshape(-1)

    # collision constraint: per-group sums of min(0, margin - score) along the densified path (optim.py:190-207) --
    def collision_tensor(self, p):
        dense = utils.dense_path(p, self.max_speed)
        slack = -(self.dist_est(dense[1:-1]) - self.safety_margin)
        self.cnt_check += len(dense)
        slack = torch.clamp(slack, max=0).reshape(-1)
        n_seg, n_pt = len(p) - 1, len(dense) - 2
        width = -(-n_pt // n_seg)
        if n_seg * width != n_pt:
            slack = torch.cat([slack, torch.zeros(n_seg * width - n_pt, dtype=slack.dtype)])
        return slack.reshape(n_seg, -1).sum(dim=1)

    def con_collision(self, x):
        with torch.no_grad():
            return self.collision_tensor(self.full_path(x).detach()).numpy()

    def jac_con_collision(self, x):
        p = self.full_path(x)
        jac = torch.autograd.functional.jacobian(self.collision_tensor, p, create_graph=False, strict=False, vectorize=True,
                                                 strategy="reverse-mode")
        return jac[:, 1:-1].numpy().reshape(jac.shape[0], -1)

    def hess_con_collision(self, x, v, fd_step=1e-3):
        """sum_k v_k Hessian(c_k)(x) for trust-constr (optim.py:380-391, where it is a double-backward through the
        reference's autograd kernel).  Here: central differences of the analytic gradient of v . c — the 2n perturbed
        paths are densified with the base path's point counts (the structure autograd would hold fixed too), scored in
        ONE dist_est launch and back-propagated in one backward.  fd_step 1e-3 suits a float32 model (gradient noise
        ~1e-6 of its maximum); a float64 model tolerates 1e-5.  Counts every evaluated point in cnt_check."""
        base = self.full_path(x).detach()
        dof = base.shape[1]
        n = base[1:-1].numel()
        v = torch.as_tensor(np.asarray(v), dtype=base.dtype)
        steps, max_step = utils.segment_steps(base, self.max_speed, None)
        seg_index = torch.repeat_interleave(torch.arange(len(base) - 1), steps)
        first = torch.cumsum(steps, 0) - steps
        k = (torch.arange(len(seg_index)) - first[seg_index]).to(base.dtype).reshape(1, -1, 1)
        paths = base.repeat(2 * n, 1, 1)
        bump = fd_step * torch.eye(n, dtype=base.dtype).reshape(n, -1, dof)
        paths[:n, 1:-1] += bump
        paths[n:, 1:-1] -= bump
        paths.requires_grad_(True)
        delta = paths[:, 1:] - paths[:, :-1]
        dist = delta.norm(dim=-1, keepdim=True)
        unit = delta * max_step / torch.where(dist > 0, dist, torch.ones_like(dist))
        pts = paths[:, :-1][:, seg_index] + k * unit[:, seg_index]  # (2n, M-1, dof): dense paths without the last waypoint
        inner = pts[:, 1:]                                          # the reference scores dense[1:-1]
        n_pt, n_seg = inner.shape[1], len(base) - 1
        slack = -(self.dist_est(inner.reshape(-1, dof)).reshape(2 * n, n_pt) - self.safety_margin)
        self.cnt_check += 2 * n * (n_pt + 2)
        slack = torch.clamp(slack, max=0)
        width = -(-n_pt // n_seg)
        if n_seg * width != n_pt:
            slack = torch.cat([slack, torch.zeros(2 * n, n_seg * width - n_pt, dtype=slack.dtype)], dim=1)
        con = slack.reshape(2 * n, n_seg, width).sum(dim=2)
        total = (con * v.to(con.dtype)).sum()
        if not total.requires_grad:
            return np.zeros((n, n))
        (grad,) = torch.autograd.grad(total, paths)
        g = grad[:, 1:-1].reshape(2 * n, n).double()
        hess = (g[:n] - g[n:]) / (2 * fd_step)
        return (0.5 * (hess + hess.T)).numpy()

    # joint limits (optim.py:220-236) ----------------------------------------------------------------------------
    def con_joint_limit(self, x):
        p = self.full_path(x)
        lim = self.robot.limits
        val = -torch.sum(torch.clamp(lim[:, 0] - p, min=0) + torch.clamp(p - lim[:, 1], min=0))
        self._limit_cache = (np.array(x, copy=True), p, val)
        return val.item()

    def grad_con_joint_limit(self, x):
        if self._limit_cache is None or not np.array_equal(x, self._limit_cache[0]):
            self.con_joint_limit(x)
        _, p, val = self._limit_cache
        p.grad = None
        if val.requires_grad:
            val.backward(retain_graph=True)
        if p.grad is None:
            return np.zeros(len(x), dtype=np.asarray(x).dtype)
        return p.grad[1:-1].numpy().reshape(-1)


def givengrad_traj_optimize(robot, dist_est, start_cfg, target_cfg, options):
    """optim.py:166-321.  SLSQP on the path length with the collision constraint (and its full Jacobian) and the joint
    limits; random restarts until scipy reports success, otherwise the least-violating result."""
    n_waypoints, n_trials, maxiter = options["N_WAYPOINTS"], options["NUM_RE_TRIALS"], options["MAXITER"]
    safety_margin, max_speed = options["safety_margin"], options["max_speed"]
    seed = options["seed"]
    torch.manual_seed(seed)
    t0 = time.time()
    cnt_check, success, best, best_violation, best_problem = 0, False, None, np.inf, None
    for trial in range(n_trials):
        path, trivial = _initial_path(robot, start_cfg, target_cfg, n_waypoints, trial, options)
        if trivial:
            return _trivial_record(robot, path, start_cfg, target_cfg, seed, t0)
        prob = _SlsqpProblem(robot, dist_est, path, safety_margin, max_speed)
        res = minimize(prob.cost, path[1:-1].reshape(-1).numpy(), jac=prob.grad_cost, method="slsqp",
                       constraints=[{"fun": prob.con_collision, "type": "ineq", "jac": prob.jac_con_collision},
                                    {"fun": prob.con_joint_limit, "type": "ineq", "jac": prob.grad_con_joint_limit}],
                       options={"maxiter": maxiter, **options.get("extra_optimizer_options", {})})
        if res.success:
            success, best, best_problem = True, res, prob
            cnt_check += prob.cnt_check
            break
        violation = -(prob.con_collision(res.x).sum() + prob.con_joint_limit(res.x))
        cnt_check += prob.cnt_check
        if violation < best_violation:
            best_violation, best, best_problem = violation, res, prob
    solution = best_problem.full_path(best.x).detach()
    return {"start_cfg": _cfg_list(start_cfg), "target_cfg": _cfg_list(target_cfg), "cnt_check": cnt_check,
            "cost": float(best.fun), "time": time.time() - t0, "success": success, "seed": seed,
            "solution": _cfg_list(solution)}


def trustconstr_traj_optimize(robot, dist_est, start_cfg, target_cfg, options):
    """optim.py:324-507.  scipy trust-constr on the path length with the collision constraint (Jacobian + Hessian) and the
    joint limits; random restarts until scipy reports success, otherwise the least-violating result.  The record carries
    scipy's result under ``info`` like the reference's.  ``options['hess_fd_step']`` (default 1e-3) is the step of the
    finite-difference Hessian."""
    n_waypoints, n_trials, maxiter = options["N_WAYPOINTS"], options["NUM_RE_TRIALS"], options["MAXITER"]
    safety_margin, max_speed = options["safety_margin"], options["max_speed"]
    fd_step = options.get("hess_fd_step", 1e-3)
    seed = options["seed"]
    torch.manual_seed(seed)
    t0 = time.time()
    cnt_check, success, best, best_violation, best_problem = 0, False, None, np.inf, None
    for trial in range(n_trials):
        path, trivial = _initial_path(robot, start_cfg, target_cfg, n_waypoints, trial, options)
        if trivial:
            return _trivial_record(robot, path, start_cfg, target_cfg, seed, t0)
        prob = _SlsqpProblem(robot, dist_est, path, safety_margin, max_speed)
        res = minimize(prob.cost, path[1:-1].reshape(-1).numpy(), jac=prob.grad_cost, method="trust-constr",
                       constraints=[NonlinearConstraint(prob.con_collision, 0, np.inf, jac=prob.jac_con_collision,
                                                        hess=lambda x, v: prob.hess_con_collision(x, v, fd_step)),
                                    NonlinearConstraint(prob.con_joint_limit, 0, np.inf, jac=prob.grad_con_joint_limit)],
                       options={"maxiter": maxiter, **options.get("extra_optimizer_options", {})})
        if res.success:
            success, best, best_problem = True, res, prob
            cnt_check += prob.cnt_check
            break
        violation = -(prob.con_collision(res.x).sum() + prob.con_joint_limit(res.x))
        cnt_check += prob.cnt_check
        if violation < best_violation:
            best_violation, best, best_problem = violation, res, prob
    solution = best_problem.full_path(best.x).detach()
    return {"start_cfg": _cfg_list(start_cfg), "target_cfg": _cfg_list(target_cfg), "cnt_check": cnt_check,
            "cost": float(best.fun), "time": time.time() - t0, "success": success, "seed": seed,
            "solution": _cfg_list(solution), "info": best}


def gradient_free_traj_optimize(robot, dist_est, start_cfg, target_cfg, options):
    raise NotImplementedError("the gradient-free baseline (optim.py:519-629) exists to time a geometric checker; it is not "
                              "part of the differentiable hot path")


class TrajOptimizer:
    """optim.py:632-661."""

    def __init__(self, robot, checker, options):
        self.robot, self.checker, self.options = robot, checker, options
        self.normalizer = lambda x: x
        self.unnormalizer = lambda x: x

    def step(self, x):
        raise NotImplementedError

    def set_unnormalizer(self, f):
        self.unnormalizer = f

    def set_normalizer(self, f):
        self.normalizer = f

    def set_checker(self, checker):
        self.checker = checker

    def set_robot(self, robot):
        self.robot = robot


class Weighted(TrajOptimizer):
    """optim.py:664-761 — the device-resident penalty optimiser used by the active-learning scripts: Adam (or any torch
    optimiser) on the waypoints, collision hinge on ``checker.rbf_score`` (optionally along the densified path),
    early exit once the constraint loss drops to 0.5."""

    def __init__(self, robot, checker, options):
        super().__init__(robot, checker, options)
        self.n_waypoints = options["n_waypoints"]
        self.maxiter = options["maxiter"]
        self.history = options["history"]
        self.dif_weight = 1
        self.max_move_weight = options["max_move_weight"]
        self.collision_weight = options["collision_weight"]
        self.joint_limit_weight = options["joint_limit_weight"]
        self.safety_bias = options["safety_bias"]
        self.max_speed = options["max_speed"]
        self.optimizer = options["optimizer"]
        self.optimizer_params = options["optimizer_params"]
        self.dense_check = options["dense_check"]
        self.fused = bool(options.get("fused", False))
        self._fused_cache = None  # (key, GraphedWeightedStep): the captured CUDA graph survives across step() calls
        self._logger = None

    def setup_logger(self, logger):
        self._logger = logger

    @torch.inference_mode(False)
    def step(self, p, maxiter=None, mask=None, write=True, verbose=False):
        assert torch.is_grad_enabled() and not torch.is_inference_mode_enabled()
        t0 = time.time()
        if not isinstance(p, torch.Tensor):
            p = torch.as_tensor(np.asarray(p), dtype=torch.float32)
        p = self.unnormalizer(p).to(self.checker.device)
        maxiter = self.maxiter if maxiter is None else maxiter
        if self.fused:
            from . import trajopt

            p, history = trajopt.fused_weighted_steps(self, p, maxiter, mask, verbose)
        else:
            p, history = self._autograd_steps(p, maxiter, mask, verbose)
        out = self.normalizer(p.detach().cpu())
        return OptimizerResult(x=out, misc={"path_history": history, "time": time.time() - t0})

    def _autograd_steps(self, p, maxiter, mask, verbose):
        p = p.detach().clone().requires_grad_(True)
        opt = self.optimizer([p], **self.optimizer_params)
        dist_est = self.checker.rbf_score
        limits = self.robot.limits.to(p.device)
        history = []
        for step in range(maxiter):
            opt.zero_grad()
            collision = 0
            if self.collision_weight != 0:
                check_p = utils.dense_path(p, max_step=self.max_speed) if self.dense_check else p
                collision = torch.clamp(dist_est(check_p) + self.safety_bias, min=0).mean() * len(p)
            cp = self.robot.fkine(p)
            seg = (cp[1:] - cp[:-1]).square()
            max_move = torch.clamp(seg.sum(dim=2) - self.max_speed**2, min=0).sum() if self.max_move_weight != 0 else 0
            joint_limit = 0
            if self.joint_limit_weight != 0:
                joint_limit = (torch.clamp(limits[:, 0] - p, min=0) + torch.clamp(p - limits[:, 1], min=0)).sum()
            constraint = (self.collision_weight * collision + self.max_move_weight * max_move
                          + self.joint_limit_weight * joint_limit)
            loss = self.dif_weight * seg.sum() + constraint
            loss.backward()
            if mask is not None:
                p.grad[~mask] = 0.0
            opt.step()
            p.data = self.robot.wrap(p.data)
            if verbose and self._logger is not None and (step % max(1, maxiter // 5) == 0 or step + 1 == maxiter):
                self._logger.info(f"obj {float(seg.sum()):.3f}x1, col {float(collision):.3f}x{self.collision_weight}, "
                                  f"jnt {float(joint_limit):.3f}x{self.joint_limit_weight}, "
                                  f"spd {float(max_move):.3f}x{self.max_move_weight}.")
            if self.history:
                history.append(self.normalizer(p.detach().cpu()))
            if float(constraint.detach()) <= 0.5:
                break
        return p, history
