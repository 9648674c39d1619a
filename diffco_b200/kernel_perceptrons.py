"""Kernel perceptrons — host-side mirror of the reference's ``diffco/kernel_perceptrons.py`` (``DiffCo``) and of the
working legacy multi-class perceptron ``diffco/deprecated/MultiDiffCo.py`` (``MultiDiffCo``).

Same constructors, methods, attribute names and return shapes as the reference, so ``dist_est = checker.poly_score``
(or ``.score`` / ``.rbf_score``) drops into the reference's unmodified optimisers (diffco/optim.py).  All arithmetic is
native: scoring and its gradient are one fused CUDA launch (``dc_score_grad``), training is one persistent CUDA
launch (``dc_perceptron_train``), kernel matrices come from ``dc_kernel_matrix``; only ``torch.linalg.solve`` in
``fit_poly`` is a library call, as in the reference (kernel_perceptrons.py:283).

Device semantics: model tensors live on the CUDA device; queries may live anywhere (the reference optimisers pass CPU
float64 waypoints) and results are returned on the query's device, so CPU-side callers keep working unchanged.
"""
from __future__ import annotations

import ctypes as C
from time import time


import torch

from . import _lib, functional, kernel
from ._lib import DC_GRAD_JAC, DC_GRAD_NONE, DC_GRAD_SUM


def _resolve_transform(transform):
    """Accept what the reference accepts for ``transform`` — a robot's bound ``fkine`` (or the robot itself) — and
    return (callable, dc_fk_desc).  Arbitrary Python callables cannot be fused into the CUDA kernel."""
    if transform is None:
        return None, None
    owner = getattr(transform, "__self__", None)
    if owner is not None and getattr(owner, "fk_desc", None) is not None:
        return transform, owner.fk_desc
    if getattr(transform, "fk_desc", None) is not None:
        return transform.fkine, transform.fk_desc
    raise TypeError("transform must be a diffco_b200.model robot (or its bound .fkine); arbitrary Python feature maps "
                    "cannot be fused into the CUDA score kernel")


class Perceptron:
    """kernel_perceptrons.py:12-27."""

    def __init__(self):
        self.support_points = None

    def score(self, point):
        raise NotImplementedError

    def is_collision(self, point):
        return self.score(point) > 0

    def line_predict(self, start, target, res=50):
        points = map(lambda i: start + (target - start) / res * i, range(res))
        return any(map(lambda p: self.is_collision(p), points))


class _SupportCache:
    """Packed support tables keyed on the identity + version of the (features, weights) tensors they were built from,
    so that user code assigning ``checker.gains = ...`` or editing tensors in place is picked up."""

    def __init__(self):
        self._entries = {}

    @staticmethod
    def _key(t):
        return (id(t), t._version, tuple(t.shape), t.dtype, t.device)

    def get(self, name, feats, weights, kfun=None, lo_fn=None):
        kdesc = getattr(kfun, "desc", None)
        kkey = None if kdesc is None else (kdesc.kind, kdesc.order, kdesc.param, kdesc.param2, kdesc.alpha, kdesc.order2)
        key = (self._key(feats), self._key(weights), kkey)
        hit = self._entries.get(name)
        if hit is None or hit[0] != key:
            hit = (key, functional.SupportSet(feats, weights, kernel=kdesc, s_lo=(lo_fn() if lo_fn is not None else None)))
            self._entries[name] = hit
        return hit[1]

    def clear(self):
        self._entries.clear()


class _FusedScorer:
    """Shared machinery: evaluate sum_n w k(FK(q), s_n) (+ Jacobian) for queries living on any device."""

    def _fk_for(self, sv):
        return self._fk_desc if self._fk_desc is not None else functional.none_fk(sv.n_features)

    def _evaluator(self, sv, kdesc, fk):
        def run(q_in, mode, grad_out=None):
            """mode DC_GRAD_NONE -> (score, None); DC_GRAD_JAC -> (score, (B, C, D)); DC_GRAD_SUM -> (score, (B, D)) =
            d(sum_c grad_out[:, c] score[:, c])/dq."""
            q = q_in.detach().to(device=sv.device, dtype=sv.dtype)
            if grad_out is not None:
                grad_out = grad_out.detach().to(device=sv.device, dtype=sv.dtype)
            score, grad = functional.score_grad(fk, kdesc, sv, q, mode, grad_out)
            score = score.to(q_in.device)
            if grad is not None:
                grad = grad.to(device=q_in.device, dtype=q_in.dtype)
            return score, grad

        run.n_class = sv.n_class
        return run

    def _fused(self, point, sv, kernel_func, fk):
        """(B, D) query -> (B, C) scores, differentiable w.r.t. ``point``."""
        return functional.differentiable_score(point, self._evaluator(sv, kernel_func.desc, fk))

    def score_and_grad(self, point, weights="gains", grad_out=None):
        """Fast path for callers that know the upstream gradient in advance (e.g. differentiating ``score.sum()``):
        one launch returning ``score (B, C)`` and ``d(sum_c grad_out[:,c] score[:,c])/dq (B, D)``, no autograd graph.
        ``weights`` selects ``gains`` (score) or ``rbf_nodes`` (poly_score / rbf_score)."""
        sv, kfun = self._select(weights)
        q = point if point.ndim == 2 else point[None, :]
        qd = q.detach().to(device=sv.device, dtype=sv.dtype)
        score, grad = functional.score_grad(self._fk_for(sv), kfun.desc, sv, qd, DC_GRAD_SUM, grad_out)
        return score.to(q.device), grad.to(q.device)


class DiffCo(Perceptron, _FusedScorer):
    """kernel_perceptrons.py:31-370.  Also accepts the legacy constructor ``DiffCo(obstacles, kernel_func=FKKernel(...),
    beta=...)`` (deprecated/DiffCo.py:30-38) that the reference's scripts use."""

    _NEW_ARGS = ("kernel_func", "gamma", "beta", "transform", "max_batch_size", "max_num_supports")
    _LEGACY_ARGS = ("obstacles", "kernel_func", "gamma", "beta", "gt_checker")

    def __init__(self, *args, **kwargs):
        super().__init__()
        legacy = "obstacles" in kwargs or "gt_checker" in kwargs or (
            len(args) > 0 and not isinstance(args[0], (str, kernel.KernelFunc)))
        names = self._LEGACY_ARGS if legacy else self._NEW_ARGS
        if len(args) > len(names):
            raise TypeError(f"DiffCo takes at most {len(names)} positional arguments")
        params = dict(zip(names, args))
        for k, v in kwargs.items():
            if k in params:
                raise TypeError(f"DiffCo got multiple values for argument {k!r}")
            if k not in self._NEW_ARGS + self._LEGACY_ARGS:
                raise TypeError(f"DiffCo got an unexpected keyword argument {k!r}")
            params[k] = v
        kernel_func = params.get("kernel_func", "rq")
        gamma, beta = params.get("gamma", 1), params.get("beta", 1)
        transform = params.get("transform")
        max_batch_size, max_num_supports = params.get("max_batch_size"), params.get("max_num_supports")
        self.obstacles = params.get("obstacles")
        self.gt_checker = params.get("gt_checker")
        if max_num_supports is not None:
            raise NotImplementedError("max_num_supports: only the max_num_supports=None path is implemented "
                                      "(the reference's truncation keeps the smallest gains, kernel_perceptrons.py:175)")
        self.train_method = None
        self.kernel_func = kernel.RQKernel(gamma) if isinstance(kernel_func, str) and kernel_func == "rq" else kernel_func
        if isinstance(self.kernel_func, kernel.FKKernel):
            transform = self.kernel_func.fkine
            self.kernel_func = self.kernel_func.rq_kernel
        elif isinstance(self.kernel_func, (kernel.LineFKKernel, kernel.TemporalFKKernel)):
            # kernel.py:145-202: rows are [q_a | q_b] segments / [q | t] space-time points; the composite feature map and
            # the radial part are fused into the score kernel like a plain robot + RQ kernel
            transform = self.kernel_func.map.fkine
            self.kernel_func = self.kernel_func.radial
        if not isinstance(self.kernel_func, kernel.KernelFunc) or self.kernel_func.desc is None:
            raise TypeError("kernel_func must be 'rq' or a diffco_b200.kernel radial kernel")
        self.beta = beta
        self.transform, self._fk_desc = _resolve_transform(transform)
        self.fkine = self.transform
        self._cuda = True

        self.support_points = None
        self.support_transformed = None
        self.gains = None
        self.hypothesis = None
        self.y = None
        self.distance = None
        self.kernel_matrix = None
        self.rbf_nodes = None
        self.rbf_kernel = None
        self.max_batch_size = max_batch_size
        self.max_num_supports = None
        self._valid_supports = 0
        self.train_iterations = None
        self._cache = _SupportCache()

    # ------------------------------------------------------------------ training
    def _features(self, X):
        """transform(X) flattened, on the GPU, through dc_fk_forward."""
        if self._fk_desc is None:
            return X.reshape(X.shape[0], -1)
        return functional.fk_forward(self._fk_desc, X)

    def _shape_features(self, Xf):
        if self._fk_desc is None:
            return Xf
        return Xf.reshape(Xf.shape[0], self._fk_desc.n_points, self._fk_desc.point_dim)

    def train(self, X, y, update=False, exist_mask=None, max_iteration=1000, method="original", distance=None,
              verbose=False, keep_all=False):
        """kernel_perceptrons.py:56-80 -> train_perceptron :98-158 (greedy loop on the device)."""
        if method != "original":
            raise NotImplementedError(f"train method {method!r} (the reference implements only 'original')")
        dev = functional._require_cuda()
        self.train_method = method
        dtype = X.dtype if X.dtype in (torch.float32, torch.float64) else torch.float32
        X = X.detach().to(device=dev, dtype=dtype)
        y = y.detach().to(device=dev, dtype=dtype).reshape(-1)
        assert len(y) == len(X)
        self.distance = distance.detach().to(dev).reshape(-1) if distance is not None else None
        t0 = time()
        n = len(X)
        # Only the kernel rows the greedy loop asks for are stored (the reference zero-initialises the full N x N matrix,
        # kernel_perceptrons.py:90-96: 400 MB at N = 10 000): rows[cap, N] + a slot map; a run that needs more rows than
        # `cap` reports it and is repeated with cap = N (never observed with the default: supports are a small fraction).
        lib = _lib.load()
        cap = n if getattr(self, "full_kernel_rows", False) else min(n, max(2048, n // 3))
        while True:
            if update:
                gains, Xf, hyp, rows0, slot = self._jump_start(X, exist_mask.to(dev))
                cap = max(cap, len(rows0))
                K = torch.zeros((cap, n), dtype=dtype, device=dev)
                K[:len(rows0)] = rows0
                diag = torch.zeros(n, dtype=dtype, device=dev)
                e = torch.where(slot >= 0)[0]
                diag[e] = rows0[slot[e].long(), e]
            else:
                Xf = self._features(X).contiguous()
                gains = torch.zeros(n, dtype=dtype, device=dev)
                hyp = torch.zeros(n, dtype=dtype, device=dev)
                K = torch.zeros((cap, n), dtype=dtype, device=dev)
                slot = torch.full((n,), -1, dtype=torch.int32, device=dev)
                diag = torch.zeros(n, dtype=dtype, device=dev)
            iters = torch.zeros(2, dtype=torch.int64, device=dev)
            with torch.cuda.device(dev):
                st = lib.dc_perceptron_train_rows(C.byref(self.kernel_func.desc), Xf.data_ptr(), y.data_ptr(), n, Xf.shape[1], 1,
                                                  functional._dtype_code(dtype), float(self.beta), int(max_iteration),
                                                  gains.data_ptr(), hyp.data_ptr(), K.data_ptr(), slot.data_ptr(), cap,
                                                  diag.data_ptr(), 0, iters.data_ptr(), functional._stream_ptr(dev))
            _lib.check(st, "dc_perceptron_train_rows")
            it_last, rows_done = (int(v) for v in iters.tolist())
            if rows_done >= 0 or cap == n:
                break
            cap = n  # out of row storage: once more with room for every row
        self.train_iterations = it_last
        self.train_kernel_rows = rows_done
        if verbose:
            print(f"Ended at iteration {self.train_iterations}, cost {time() - t0:.4f} secs")
            print("ACC: {}".format(torch.sum((hyp > 0) == (y > 0)) / float(len(y))))

        mask = gains != 0
        if keep_all:
            mask = torch.ones_like(mask)
        if mask.sum() < 2:  # kernel_perceptrons.py:140-141
            mask[torch.where(mask == 0)[0][0]] = True
        idx = torch.where(mask)[0]
        self.support_index = idx
        self.support_points = X[mask]
        self.support_transformed = self._shape_features(Xf[mask])
        self.hypothesis = hyp[mask]
        self.y = y[mask]
        self.distance = self.distance[mask] if self.distance is not None else None
        self.gains = gains[mask]
        self.rbf_nodes = self.gains.new_zeros(len(self.gains))
        # every support had its row computed, except a never-updated point kept only to have two supports (:140-141)
        srow = slot[idx].long()
        self.kernel_matrix = K[srow.clamp_min(0)][:, idx]
        missing = srow < 0
        if bool(missing.any()):
            fill = functional.kernel_matrix(self.kernel_func.desc, Xf[idx[missing]], Xf[idx])
            self.kernel_matrix[missing] = fill
            self.kernel_matrix[:, missing] = fill.T
        self._valid_supports = len(self.support_points)
        self._cache.clear()
        if verbose:
            print(f"DiffCo training done. {time() - t0:.4f} secs cost")

    def _jump_start(self, X, exist_mask):
        """jump_start_initialize, kernel_perceptrons.py:222-269: warm start from the current supports."""
        n = len(X)
        dev, dtype = X.device, X.dtype
        novel = X[~exist_mask]
        assert n - len(novel) == self.valid_supports
        hyp = torch.zeros(n, dtype=dtype, device=dev)
        hyp[exist_mask] = self.hypothesis.to(dtype)
        hyp[~exist_mask] = self.score_original(novel).reshape(-1).to(dtype)
        novel_f = self._features(novel)
        sup_f = self.support_transformed.reshape(self.valid_supports, -1).to(dtype)
        e = torch.where(exist_mask)[0]
        v = torch.where(~exist_mask)[0]
        # kernel rows of the existing supports against ALL points (old block + cross block); the novel points' rows are
        # computed by the training loop when it asks for them.  Row k belongs to the k-th existing point.
        rows = torch.zeros((len(e), n), dtype=dtype, device=dev)
        rows[:, e] = self.kernel_matrix.to(dtype)
        rows[:, v] = functional.kernel_matrix(self.kernel_func.desc, sup_f, novel_f)
        slot = torch.full((n,), -1, dtype=torch.int32, device=dev)
        slot[e] = torch.arange(len(e), dtype=torch.int32, device=dev)
        Xf = torch.zeros((n, sup_f.shape[1]), dtype=dtype, device=dev)
        Xf[exist_mask] = sup_f
        Xf[~exist_mask] = novel_f
        gains = torch.zeros(n, dtype=dtype, device=dev)
        gains[exist_mask] = self.gains.to(dtype)
        check = rows.T @ gains[e]  # K @ gains with gains == 0 outside the existing supports
        assert torch.allclose(check, hyp, atol=1e-4), f"diff: {torch.abs(check - hyp).max()}"  # kernel_perceptrons.py:266-268
        return gains, Xf.contiguous(), hyp, rows, slot

    @property
    def valid_supports(self):
        return self._valid_supports

    def fit_poly(self, kernel_func, target="hypo", reg=0.0):
        """kernel_perceptrons.py:271-287: rbf_nodes = solve(rbf_kernel(S, S), target)."""
        if target == "hypo":
            t = self.hypothesis
        elif "dist" in target:
            t = self.distance
        elif "label" in target:
            t = self.y
        else:
            raise ValueError(target)
        self.rbf_kernel = kernel_func
        S = self.support_transformed.reshape(self.valid_supports, -1)
        kmat = functional.kernel_matrix(kernel_func.desc, S, S)
        if reg:
            kmat = kmat + reg * torch.eye(len(kmat), dtype=kmat.dtype, device=kmat.device)
        self.rbf_nodes = torch.linalg.solve(kmat, t[:, None].to(kmat.dtype)).reshape(-1)
        self._cache.clear()

    # ------------------------------------------------------------------ device management (kernel_perceptrons.py:289-307)
    def cuda(self):
        return self.to(functional._require_cuda())

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("diffco_b200 perceptrons live on the GPU; queries may be CPU tensors, the model may not")
        for name in ("support_points", "support_transformed", "gains", "rbf_nodes", "hypothesis", "y", "distance", "kernel_matrix"):
            t = getattr(self, name, None)
            if t is not None:
                setattr(self, name, t.to(device))  # the reference forgets gains here (kernel_perceptrons.py:297-303)
        self._cache.clear()
        return self

    @property
    def device(self):
        return self.support_points.device

    def __getstate__(self):
        state = self.__dict__.copy()
        state["_cache"] = None
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._cache = _SupportCache()

    # ------------------------------------------------------------------ scoring
    def _select(self, weights):
        if weights == "gains":
            return self._cache.get("gains", self.support_transformed, self.gains, self.kernel_func, self._support_lo), self.kernel_func
        return self._cache.get("rbf", self.support_transformed, self.rbf_nodes, self.rbf_kernel, self._support_lo), self.rbf_kernel

    def _support_lo(self):
        """Low parts of the support features (what float32 rounding dropped from FK(support_points)), for the tensor-core
        kernel's exact near-pair path — available when the float32 ``support_transformed`` IS this package's FK of
        ``support_points`` (checked bit for bit; anything else, e.g. user-assigned features, simply goes without)."""
        sp, st = self.support_points, self.support_transformed
        if self._fk_desc is None or sp is None or st is None or st.dtype != torch.float32 or not st.is_cuda:
            return None
        if sp.ndim != 2 or sp.shape[0] != st.shape[0] or sp.shape[1] != self._fk_desc.dof:
            return None
        hi, lo = functional.fk_forward_split(self._fk_desc, sp.to(device=st.device, dtype=torch.float32))
        return lo if torch.equal(hi, st.reshape(st.shape[0], -1)) else None

    def score(self, point):
        return self.score_original(point)

    def score_original(self, point):
        """kernel_perceptrons.py:362-370 — (B,), or 0-dim when B == 1 and the kernel squeezes (kernel.py:26-27)."""
        if point.ndim == 1:
            point = point[None, :]
        sv, kfun = self._select("gains")
        s = self._fused(point, sv, kfun, self._fk_for(sv)).reshape(-1)
        if s.shape[0] == 1 and kfun.squeeze_single_row:
            s = s.reshape(())
        return s

    def poly_score(self, point=None, transformed_point=None):
        """kernel_perceptrons.py:309-319 — (B, 1); input is cast to the model's dtype (:313)."""
        sv, kfun = self._select("rbf")
        if transformed_point is None:
            if point.ndim == 1:
                point = point.unsqueeze(0)
            s = self._fused(point, sv, kfun, self._fk_for(sv))
        else:
            x = transformed_point.reshape(transformed_point.shape[0], -1)
            s = self._fused(x, sv, kfun, functional.none_fk(sv.n_features))
        if s.shape[0] == 1 and kfun.squeeze_single_row:
            s = s.reshape(1)
        return s

    rbf_score = poly_score  # legacy name (deprecated/DiffCo.py, kernel_perceptrons.py:532-540)


class MultiDiffCo(DiffCo):
    """Legacy multi-class perceptron, diffco/deprecated/MultiDiffCo.py:19-170 (the package-level ``MultiDiffCo`` of the
    reference is stale).  gains / hypothesis / rbf_nodes are (N, C); one kernel matrix is shared by
    all classes; ``support_points`` are raw configurations and the kernel is ``FKKernel(fkine, radial)``."""

    def __init__(self, objects=None, kernel_func="rq", gamma=1, beta=1, gt_checker=None, transform=None):
        super().__init__(kernel_func=kernel_func, gamma=gamma, beta=beta, transform=transform)
        self.objects = objects
        self.obstacles = objects
        self.gt_checker = gt_checker
        self.num_class = None
        self.support_fkine = None

    def train(self, X, y, max_iteration=1000, gains=None, hypothesis=None, method="original", distance=None,
              kernel_matrix=None):
        """deprecated/MultiDiffCo.py:23-83."""
        if method != "original":
            raise NotImplementedError(method)
        dev = functional._require_cuda()
        dtype = X.dtype if X.dtype in (torch.float32, torch.float64) else torch.float32
        X = X.detach().to(device=dev, dtype=dtype)
        Y = y.detach().to(device=dev, dtype=dtype).reshape(len(X), -1).contiguous()
        n, self.num_class = Y.shape
        self.train_method, self.distance = method, distance
        if gains is None and hypothesis is None and kernel_matrix is None:
            G = torch.zeros((n, self.num_class), dtype=dtype, device=dev)
            H = torch.zeros((n, self.num_class), dtype=dtype, device=dev)
            K = torch.zeros((n, n), dtype=dtype, device=dev)
        elif gains is None or hypothesis is None or kernel_matrix is None:
            raise ValueError("DiffCo: you passed in some existing parameters but not all three of gains, hypothesis, and kernel_matrix")
        else:
            G, H, K = (t.detach().to(device=dev, dtype=dtype).contiguous().clone() for t in (gains, hypothesis, kernel_matrix))
        Xf = self._features(X).contiguous()
        diag = torch.diagonal(K).clone()
        iters = torch.zeros(2, dtype=torch.int64, device=dev)
        lib = _lib.load()
        with torch.cuda.device(dev):
            st = lib.dc_perceptron_train(C.byref(self.kernel_func.desc), Xf.data_ptr(), Y.data_ptr(), n, Xf.shape[1],
                                         self.num_class, functional._dtype_code(dtype), float(self.beta), int(max_iteration),
                                         G.data_ptr(), H.data_ptr(), K.data_ptr(), diag.data_ptr(), 1, iters.data_ptr(),
                                         functional._stream_ptr(dev))
        _lib.check(st, "dc_perceptron_train")
        self.train_iterations = int(iters[0].item())
        keep = torch.sum(G != 0, dim=1) != 0  # deprecated/MultiDiffCo.py:34-35
        idx = torch.where(keep)[0]
        self.support_index = idx
        self.support_points = X[keep]
        self.support_transformed = self._shape_features(Xf[keep])
        self.support_fkine = Xf[keep]
        self.hypothesis, self.y, self.gains = H[keep], Y[keep], G[keep]
        self.distance = self.distance[keep] if self.distance is not None else None
        self.kernel_matrix = K[idx[:, None], idx[None, :]]
        self.rbf_nodes = None
        self._valid_supports = len(idx)
        self._cache.clear()

    def fit_poly(self, kernel_func=None, target="hypo", fkine=None, reg=0):
        """deprecated/MultiDiffCo.py:125-154: per class, kernel entries coupling a support used by the class with one
        unused by it are zeroed; solve; nodes are zeroed where the gain is zero."""
        if fkine is not None:
            self.transform, self._fk_desc = _resolve_transform(fkine)
            self.fkine = self.transform
            self.support_fkine = self._features(self.support_points)
            self.support_transformed = self._shape_features(self.support_fkine)
        if target == "hypo":
            t = self.hypothesis
        elif "dist" in target:
            t = self.distance
        elif "label" in target:
            t = self.y
        else:
            raise ValueError(target)
        self.rbf_kernel = kernel.MultiQuadratic(1) if kernel_func is None else kernel_func
        S = self.support_transformed.reshape(len(self.support_points), -1)
        kmat = functional.kernel_matrix(self.rbf_kernel.desc, S, S)
        for c in range(self.num_class):
            nz = self.gains[:, c] != 0
            cut = nz[:, None] & (~nz)[None, :]
            kmat[cut | cut.T] = 0
        nodes = torch.linalg.solve(kmat + reg * torch.eye(len(kmat), dtype=kmat.dtype, device=kmat.device), t.to(kmat.dtype))
        nodes[self.gains == 0] = 0
        self.rbf_nodes = nodes
        assert self.rbf_nodes.shape == (len(self.support_points), self.num_class)
        self._cache.clear()

    def predict(self, point):
        return (self.score(point) > 0) * 2 - 1

    def __call__(self, *args, **kwargs):
        return self.predict(*args, **kwargs)

    def score(self, points):
        """deprecated/MultiDiffCo.py:118-123 — (B, C) ((C,) when B == 1 and the kernel squeezes)."""
        if points.ndim == 1:
            points = points[None, :]
        sv, kfun = self._select("gains")
        s = self._fused(points, sv, kfun, self._fk_for(sv))
        return s.reshape(-1) if (s.shape[0] == 1 and kfun.squeeze_single_row) else s

    score_original = score

    def rbf_score(self, point):
        """deprecated/MultiDiffCo.py:156-170 — (B, C)."""
        if point.ndim == 1:
            point = point[None, :]
        sv, kfun = self._select("rbf")
        s = self._fused(point, sv, kfun, self._fk_for(sv))
        return s.reshape(-1) if (s.shape[0] == 1 and kfun.squeeze_single_row) else s

    poly_score = rbf_score
