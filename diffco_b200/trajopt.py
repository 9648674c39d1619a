"""Device-resident fast path of ``Weighted.step`` (reference ``diffco/optim.py:686-761``).

The reference evaluates the penalty (collision hinge on ``checker.rbf_score`` + max-move + joint-limit + path length)
with ~40 small tensor ops per step and differentiates it with autograd.  Here one step is

    dc_traj_dense_path  (options['dense_check'] only: utils.dense_path on the device, utils.py:87-102)
    dc_score_grad       (scores of the W waypoints — or of the dense points — and their analytic gradient, one launch)
    dc_traj_step_ex     (control points, path-length / max-move / joint-limit / collision terms, the analytic gradient of
                         their weighted sum through J_FK^T (and through the interpolation of the dense points), the mask,
                         torch.optim.Adam's update, robot.wrap and the reference's exit test: one launch, csrc/dc_traj.cu)

with NO autograd graph, captured ONCE in a CUDA graph (cached on the optimiser across ``step()`` calls) and replayed per
iteration.  The reference's early exit (``constraint_loss <= 0.5``, optim.py:747-752) is evaluated on the device — once it
fires, further replays are no-ops — so the host reads the state back only every few iterations unless it records the
path history.  The arithmetic is the
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
                 joint_limit_weight=10.0, safety_bias=0.0, max_speed=1.0, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, wrap=True,
                 dense_check=False, max_dense_points=None, exit_constraint=-1.0):
        self.lib = _lib.load()
        self.dense_check = bool(dense_check) and float(collision_weight) != 0
        self.exit_constraint = float(exit_constraint)
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
        self.state = torch.zeros(2, device=dev, dtype=torch.int32)  # [exit test fired, iterations applied]
        if self.dense_check:
            # static upper bound on the number of dense points (the graph's scoring launch has this many rows); a path that
            # needs more is reported through `dense_count` = -1 and the caller falls back to the autograd step
            self.max_points = int(max_dense_points or max(1024, 16 * self.W))
            self.dense = torch.zeros((self.max_points, self.D), device=dev, dtype=dt)
            self.seg_offset = torch.zeros(self.W, device=dev, dtype=torch.int32)
            self.dense_count = torch.zeros(1, device=dev, dtype=torch.int32)
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
    def reset(self, p_new, mask=None):
        """New waypoints (same shape) and mask, fresh optimiser state; the captured graph is reused."""
        self.p.copy_(p_new.to(device=self.p.device, dtype=self.p.dtype))
        if self.mask is not None and mask is not None:
            m = torch.as_tensor(mask, device=self.p.device)
            m = m.reshape(-1, 1) if m.ndim == 1 else m
            self.mask.copy_(m.to(self.p.dtype).expand(self.W, self.D))
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self.step_count.zero_()
        self.state.zero_()

    def replay(self):
        """Runs one iteration and returns the six terms as Python floats (one host round trip)."""
        self.graph.replay()
        return self.terms.tolist()

    def replay_many(self, n):
        """Replays the iteration n times without touching the host; returns (exit test fired, iterations applied so far).
        Once the exit test has fired the remaining replays do nothing, so the waypoints are exactly those of the
        iteration that fired it."""
        for _ in range(n):
            self.graph.replay()
        fired, done = self.state.tolist()
        return bool(fired), int(done)

    def dense_overflow(self):
        return self.dense_check and int(self.dense_count.item()) < 0

    @torch.no_grad()
    def _one_step(self):
        p = self.p
        s_ptr = g_ptr = dense = None
        stream = functional._stream_ptr(p.device)
        if self.collision_weight != 0:
            pts = p
            if self.dense_check:
                with torch.cuda.device(p.device):
                    _lib.check(self.lib.dc_traj_dense_path(p.data_ptr(), self.W, self.D, self.dtype_code, float(self.prm.max_speed),
                                                           self.max_points, self.dense.data_ptr(), self.seg_offset.data_ptr(),
                                                           self.dense_count.data_ptr(), self.state.data_ptr(), stream),
                               "dc_traj_dense_path")
                pts = self.dense
            # optim.py:711-713: dist_est casts the points to the model dtype (kernel_perceptrons.py:313)
            s, g = functional.score_grad(self.fk_score, self.kfun.desc, self.sv, pts.to(self.sv.dtype), DC_GRAD_SUM)
            self._s, self._g = s.to(p.dtype).contiguous(), g.to(p.dtype).contiguous()
            if self.dense_check:
                self._dense = _lib.TrajDense(self.dense_count.data_ptr(), self.seg_offset.data_ptr(), self._s.data_ptr(),
                                             self._g.data_ptr(), self.max_points, 0)
                dense = C.byref(self._dense)
            else:
                s_ptr, g_ptr = self._s.data_ptr(), self._g.data_ptr()
        with torch.cuda.device(p.device):
            st = self.lib.dc_traj_step_ex(C.byref(self.fk_path), C.byref(self.prm), self.W, self.dtype_code, p.data_ptr(), s_ptr,
                                          g_ptr, dense, None if self.mask is None else self.mask.data_ptr(),
                                          self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), self.step_count.data_ptr(),
                                          self.terms.data_ptr(), self.state.data_ptr(), self.exit_constraint, stream)
        _lib.check(st, "dc_traj_step_ex")


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
                         betas=hp.get("betas", (0.9, 0.999)), eps=hp.get("eps", 1e-8), wrap=True,
                         dense_check=opt.dense_check, max_dense_points=opt.options.get("max_dense_points"),
                         exit_constraint=0.5)  # optim.py:747-752

    READBACK_EVERY = 8  # iterations between host read-backs of the exit state when no history / log line needs the terms

    def run(self, maxiter, verbose=False):
        o, history = self.o, []
        per_step = o.history or (verbose and o._logger is not None)
        step = 0
        while step < maxiter:
            if per_step:
                t = self.replay()
                if verbose and o._logger is not None and (step % max(1, maxiter // 5) == 0 or step + 1 == maxiter):
                    o._logger.info(f"obj {t[0]:.3f}x1, col {t[1]:.3f}x{o.collision_weight}, jnt {t[2]:.3f}x{o.joint_limit_weight}, "
                                   f"spd {t[3]:.3f}x{o.max_move_weight}.")
                if o.history:
                    history.append(o.normalizer(self.p.detach().cpu()))
                step += 1
                if t[4] <= 0.5:  # optim.py:747-752 (the device-side test has fired on the same value)
                    break
            else:
                n = min(self.READBACK_EVERY, maxiter - step)
                fired, _ = self.replay_many(n)
                step += n
                if fired:
                    break
        if self.dense_overflow():
            raise RuntimeError(f"dense_check: the path needs more than max_dense_points = {self.max_points} interpolation points")
        return self.p.detach().clone(), history


def _cache_key(opt, p, mask):
    sv, kfun = opt.checker._select("rbf")
    hp = opt.optimizer_params
    return (tuple(p.shape), p.dtype, p.device, id(sv), id(opt.robot), mask is None, opt.dense_check, opt.dif_weight,
            opt.max_move_weight, opt.collision_weight, opt.joint_limit_weight, opt.safety_bias, opt.max_speed,
            hp.get("lr", 1e-3), tuple(hp.get("betas", (0.9, 0.999))), hp.get("eps", 1e-8), opt.options.get("max_dense_points"))


def fused_weighted_steps(opt, p, maxiter, mask=None, verbose=False):
    """Entry point used by ``Weighted.step`` when ``options['fused']`` is set.  The captured graph is kept on the optimiser
    and reused by later ``step()`` calls with the same shapes, weights and support set (``id(sv)`` changes whenever the
    checker's supports / weights are re-packed), so only the first call pays the capture."""
    why = _eligible(opt, p)
    if why is not None:
        raise ValueError(f"options['fused'] = True: {why}")
    key = _cache_key(opt, p, mask)
    cached = getattr(opt, "_fused_cache", None)
    if cached is not None and cached[0] == key:
        stepper = cached[1]
        stepper.reset(p, mask)
    else:
        stepper = GraphedWeightedStep(opt, p, mask)
        opt._fused_cache = (key, stepper)
    return stepper.run(maxiter, verbose)


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
