"""Device-resident fast path of ``Weighted.step`` (reference ``diffco/optim.py:686-761``).

The reference evaluates the penalty (collision hinge on ``checker.rbf_score`` + max-move + joint-limit + path length)
with ~40 small tensor ops per step and differentiates it with autograd.  Here one step is

    dc_score_grad   (scores of the W waypoints and their analytic gradient, one launch)
    dc_traj_step    (control points, path-length / max-move / joint-limit / collision terms, the analytic gradient of
                     their weighted sum through J_FK^T, the mask, torch.optim.Adam's update and robot.wrap: one launch,
                     csrc/dc_traj.cu)

with NO autograd graph, captured ONCE in a CUDA graph and replayed per iteration; the only host round trip per step is
the scalar the reference's early exit looks at (``constraint_loss <= 0.5``, optim.py:747-752).  The arithmetic is the
reference's, term by term (same sums / clamps / Adam rule), so the waypoints follow the autograd path's trajectory
(tests/test_gpu_optim.py::test_weighted_step_graphed_*).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, functional
from ._lib import DC_GRAD_SUM, TrajParams


def _eligible(opt, p):
    """The graphed step covers what the reference's scripts use: Adam, a single-class checker whose feature map is a
    diffco_b200 descriptor, collision checked at the waypoints (``dense_check`` changes the number of scored points from
    step to step, which a static graph cannot follow)."""
    if not p.is_cuda:
        return "the waypoints must live on the checker's CUDA device"
    if opt.dense_check:
        return "options['dense_check'] must be False"
    if opt.optimizer is not torch.optim.Adam:
        return "options['optimizer'] must be torch.optim.Adam"
    extra = set(opt.optimizer_params) - {"lr", "betas", "eps"}
    if extra:
        return f"unsupported Adam options {sorted(extra)} (lr, betas, eps are)"
    if not hasattr(opt.robot, "fk_desc"):
        return "the robot must be a diffco_b200.model robot"
    sv, _ = opt.checker._select("rbf")
    if sv.n_class != 1:
        return "the checker must be single-class"
    return None


class GraphedPenaltyStep:
    """One iteration of the reference's penalty optimisers — a ``dc_score_grad`` launch and a ``dc_traj_step`` launch — as
    a replayable CUDA graph over static buffers.  ``terms`` (device, 6 values) holds, for the waypoints BEFORE the update
    of the last replay: path length, collision, joint limit, max move, constraint loss, |masked gradient|^2."""

    def __init__(self, robot, checker, weights, p, mask=None, *, dif_weight=1.0, max_move_weight=10.0, collision_weight=10.0,
                 joint_limit_weight=10.0, safety_bias=0.0, max_speed=1.0, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, wrap=True):
        self.lib = _lib.load()
        self.sv, self.kfun = checker._select(weights)
        if self.sv.n_class != 1:
            raise ValueError("the graphed step needs a single-class checker")
        self.fk_score = checker._fk_for(self.sv)
        self.fk_path = robot.fk_desc
        self.collision_weight = float(collision_weight)
        self.W, self.D = p.shape
        dev, dt = p.device, p.dtype
        self.dtype_code = functional._dtype_code(dt)
        self.p = p.detach().clone().contiguous()
        self.mask = None
        if mask is not None:
            m = torch.as_tensor(mask, device=dev)
            m = m.reshape(-1, 1) if m.ndim == 1 else m
            self.mask = m.to(dt).expand(self.W, self.D).contiguous()
        self.terms = torch.zeros(6, device=dev, dtype=dt)
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(self.p), torch.zeros_like(self.p)
        self.step_count = torch.zeros((), device=dev, dtype=torch.float64)
        prm = TrajParams()
        prm.dif_weight, prm.max_move_weight = float(dif_weight), float(max_move_weight)
        prm.collision_weight, prm.joint_limit_weight = float(collision_weight), float(joint_limit_weight)
        prm.safety_bias, prm.max_speed = float(safety_bias), float(max_speed)
        prm.lr, prm.beta1, prm.beta2, prm.eps = float(lr), float(betas[0]), float(betas[1]), float(eps)
        lim = robot.limits.to(dtype=dt).double().cpu()  # the reference compares against limits in the path's dtype
        wrapped = robot.wrap(torch.full((1, self.D), 100.0))[0] != 100.0 if wrap else torch.zeros(self.D, dtype=torch.bool)
        for i in range(self.D):
            prm.limits[i][0], prm.limits[i][1] = float(lim[i, 0]), float(lim[i, 1])
            prm.wrap[i] = int(wrapped[i])  # which coordinates robot.wrap maps to [-pi, pi)
        self.prm = prm
        # warm-up outside the capture (lazily initialised kernel attributes), then rewind
        p0 = self.p.clone()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                self._one_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        self.reset(p0)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._one_step()

    @torch.no_grad()
    def reset(self, p_new):
        """New waypoints (same shape), fresh optimiser state; the captured graph is reused."""
        self.p.copy_(p_new.to(device=self.p.device, dtype=self.p.dtype))
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self.step_count.zero_()

    def replay(self):
        """Runs one iteration and returns the six terms as Python floats (the one host round trip per step)."""
        self.graph.replay()
        return self.terms.tolist()

    @torch.no_grad()
    def _one_step(self):
        p = self.p
        s_ptr = g_ptr = None
        if self.collision_weight != 0:
            # optim.py:711-713: dist_est casts the waypoints to the model dtype (kernel_perceptrons.py:313)
            s, g = functional.score_grad(self.fk_score, self.kfun.desc, self.sv, p.to(self.sv.dtype), DC_GRAD_SUM)
            self._s, self._g = s.to(p.dtype).contiguous(), g.to(p.dtype).contiguous()
            s_ptr, g_ptr = self._s.data_ptr(), self._g.data_ptr()
        with torch.cuda.device(p.device):
            st = self.lib.dc_traj_step(C.byref(self.fk_path), C.byref(self.prm), self.W, self.dtype_code, p.data_ptr(), s_ptr,
                                       g_ptr, None if self.mask is None else self.mask.data_ptr(), self.exp_avg.data_ptr(),
                                       self.exp_avg_sq.data_ptr(), self.step_count.data_ptr(), self.terms.data_ptr(),
                                       functional._stream_ptr(p.device))
        _lib.check(st, "dc_traj_step")


class GraphedWeightedStep(GraphedPenaltyStep):
    """``Weighted.step`` (optim.py:686-761) on the graphed iteration."""

    def __init__(self, opt, p, mask=None):
        why = _eligible(opt, p)
        if why is not None:
            raise ValueError(f"options['fused'] = True: {why}")
        self.o = opt
        hp = opt.optimizer_params
        super().__init__(opt.robot, opt.checker, "rbf", p, mask, dif_weight=opt.dif_weight, max_move_weight=opt.max_move_weight,
                         collision_weight=opt.collision_weight, joint_limit_weight=opt.joint_limit_weight,
                         safety_bias=opt.safety_bias, max_speed=opt.max_speed, lr=hp.get("lr", 1e-3),
                         betas=hp.get("betas", (0.9, 0.999)), eps=hp.get("eps", 1e-8), wrap=True)

    def run(self, maxiter, verbose=False):
        o, history = self.o, []
        for step in range(maxiter):
            t = self.replay()
            if verbose and o._logger is not None and (step % max(1, maxiter // 5) == 0 or step + 1 == maxiter):
                o._logger.info(f"obj {t[0]:.3f}x1, col {t[1]:.3f}x{o.collision_weight}, jnt {t[2]:.3f}x{o.joint_limit_weight}, "
                               f"spd {t[3]:.3f}x{o.max_move_weight}.")
            if o.history:
                history.append(o.normalizer(self.p.detach().cpu()))
            if t[4] <= 0.5:  # optim.py:747-752
                break
        return self.p.detach(), history


def fused_weighted_steps(opt, p, maxiter, mask=None, verbose=False):
    """Entry point used by ``Weighted.step`` when ``options['fused']`` is set."""
    return GraphedWeightedStep(opt, p, mask).run(maxiter, verbose)


def scorer_of(dist_est):
    """(perceptron, 'gains' | 'rbf') when ``dist_est`` is a bound scoring method of a diffco_b200 perceptron — what
    ``adam_traj_optimize(..., options['fused'] = True)`` needs to put the collision term on the graphed step."""
    owner, name = getattr(dist_est, "__self__", None), getattr(dist_est, "__name__", "")
    if owner is None or not hasattr(owner, "_select") or not hasattr(owner, "_fk_for"):
        return None
    if name in ("score", "score_original"):
        return owner, "gains"
    if name in ("poly_score", "rbf_score"):
        return owner, "rbf"
    return None
