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


class GraphedWeightedStep:
    """One ``Weighted.step`` iteration — a ``dc_score_grad`` launch and a ``dc_traj_step`` launch — as a replayable CUDA
    graph over static buffers."""

    def __init__(self, opt, p, mask=None):
        why = _eligible(opt, p)
        if why is not None:
            raise ValueError(f"options['fused'] = True: {why}")
        self.o = opt
        self.lib = _lib.load()
        self.sv, self.kfun = opt.checker._select("rbf")
        self.fk_score = opt.checker._fk_for(self.sv)
        self.fk_path = opt.robot.fk_desc
        self.W, self.D = p.shape
        dev, dt = p.device, p.dtype
        self.dtype_code = functional._dtype_code(dt)
        self.p = p.detach().clone().contiguous()
        self.mask = None
        if mask is not None:
            m = torch.as_tensor(mask, device=dev)
            m = m.reshape(-1, 1) if m.ndim == 1 else m
            self.mask = m.to(dt).expand(self.W, self.D).contiguous()
        self.terms = torch.zeros(5, device=dev, dtype=dt)  # path length, collision, joint limit, max move, constraint
        self.exp_avg, self.exp_avg_sq = torch.zeros_like(self.p), torch.zeros_like(self.p)
        self.step_count = torch.zeros((), device=dev, dtype=torch.float64)
        hp = opt.optimizer_params
        b1, b2 = hp.get("betas", (0.9, 0.999))
        prm = TrajParams()
        prm.dif_weight, prm.max_move_weight = float(opt.dif_weight), float(opt.max_move_weight)
        prm.collision_weight, prm.joint_limit_weight = float(opt.collision_weight), float(opt.joint_limit_weight)
        prm.safety_bias, prm.max_speed = float(opt.safety_bias), float(opt.max_speed)
        prm.lr, prm.beta1, prm.beta2, prm.eps = float(hp.get("lr", 1e-3)), float(b1), float(b2), float(hp.get("eps", 1e-8))
        lim = opt.robot.limits.to(dtype=dt).double().cpu()  # the reference compares against limits in the path's dtype
        probe = torch.full((1, self.D), 100.0)
        wrapped = opt.robot.wrap(probe)[0] != 100.0            # which coordinates robot.wrap maps to [-pi, pi)
        for i in range(self.D):
            prm.limits[i][0], prm.limits[i][1] = float(lim[i, 0]), float(lim[i, 1])
            prm.wrap[i] = int(wrapped[i])
        self.prm = prm
        # warm-up outside the capture (lazily initialised kernel attributes), then rewind
        p0 = self.p.clone()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                self._one_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        with torch.no_grad():
            self.p.copy_(p0)
            self.exp_avg.zero_()
            self.exp_avg_sq.zero_()
            self.step_count.zero_()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._one_step()

    @torch.no_grad()
    def _one_step(self):
        o, p = self.o, self.p
        s_ptr = g_ptr = None
        if o.collision_weight != 0:
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

    def run(self, maxiter, verbose=False):
        o, history = self.o, []
        for step in range(maxiter):
            self.graph.replay()
            t = self.terms.tolist()  # the one host round trip per step: the reference's early-exit test
            c = t[4]
            if verbose and o._logger is not None and (step % max(1, maxiter // 5) == 0 or step + 1 == maxiter):
                o._logger.info(f"obj {t[0]:.3f}x1, col {t[1]:.3f}x{o.collision_weight}, jnt {t[2]:.3f}x{o.joint_limit_weight}, "
                               f"spd {t[3]:.3f}x{o.max_move_weight}.")
            if o.history:
                history.append(o.normalizer(self.p.detach().cpu()))
            if c <= 0.5:
                break
        return self.p.detach(), history


def fused_weighted_steps(opt, p, maxiter, mask=None, verbose=False):
    """Entry point used by ``Weighted.step`` when ``options['fused']`` is set."""
    return GraphedWeightedStep(opt, p, mask).run(maxiter, verbose)
