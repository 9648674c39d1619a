"""Tensor-facing wrappers over the C ABI: device buffers, streams and autograd plumbing only.

Everything numeric happens in libdiffco_b200.so; PyTorch supplies device memory (``torch.empty``), the current
CUDA stream and the autograd graph.  There is deliberately no CPU / eager fallback: without CUDA these
functions raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import DC_F32, DC_F64, DC_GRAD_JAC, DC_GRAD_NONE, DC_GRAD_SUM, FkDesc, KernelDesc, Supports

_DTYPES = {torch.float32: DC_F32, torch.float64: DC_F64}


def _require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("diffco_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback for this path")
    return torch.device("cuda", torch.cuda.current_device())


def _dtype_code(dtype: torch.dtype) -> int:
    try:
        return _DTYPES[dtype]
    except KeyError:
        raise TypeError(f"diffco_b200 supports float32 and float64, got {dtype}") from None


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def none_fk(n_features: int) -> FkDesc:
    """FK descriptor for ``transform=None`` (kernel_perceptrons.py:36): the input rows are the features."""
    d = FkDesc()
    d.type = _lib.DC_FK_NONE
    d.dof = n_features
    d.n_points = n_features
    d.point_dim = 1
    return d


class SupportSet:
    """Support vectors + weights packed for the fused kernels (row n = [-s_n | w_n], 16-byte aligned rows).

    ``s_feat`` is ``support_transformed.reshape(N, -1)``, ``weights`` is ``gains`` / ``rbf_nodes`` with shape (N,) or
    (N, C).  The packed table lives in HBM for the lifetime of the object (N*(F+C) elements, <= 0.4 MB for every
    configuration in BASELINE.json) and is re-read by every launch from L2.
    """

    def __init__(self, s_feat: torch.Tensor, weights: torch.Tensor, device: Optional[torch.device] = None,
                 kernel: Optional[KernelDesc] = None, s_lo: Optional[torch.Tensor] = None):
        device = device or _require_cuda()
        lib = _lib.load()
        s = s_feat.detach().reshape(s_feat.shape[0], -1)
        w = weights.detach()
        if w.ndim == 1:
            w = w[:, None]
        if w.shape[0] != s.shape[0]:
            raise ValueError(f"weights have {w.shape[0]} rows, supports have {s.shape[0]}")
        dtype = s.dtype
        code = _dtype_code(dtype)
        s = s.to(device=device, dtype=dtype).contiguous()
        w = w.to(device=device, dtype=dtype).contiguous()
        self.n, self.n_features = s.shape
        self.n_class = w.shape[1]
        f_pad, row = C.c_int32(), C.c_int32()
        _lib.check(lib.dc_supports_layout(self.n_features, self.n_class, code, C.byref(f_pad), C.byref(row)),
                   "dc_supports_layout")
        self.table = torch.empty((self.n, row.value), dtype=dtype, device=device)
        with torch.cuda.device(device):
            _lib.check(lib.dc_pack_supports(s.data_ptr(), w.data_ptr(), self.n, self.n_features, self.n_class, code,
                                            self.table.data_ptr(), _stream_ptr(device)), "dc_pack_supports")
        self.dtype = dtype
        self.device = device
        # Optional tensor-core operand image (fp32, one class, F <= 30, RQKernel with p = 2; csrc/dc_score_tc.cuh): the
        # kernel width and the weights are folded into the operands, so it is built for the kernel this support set is
        # scored with.  max|s|^2 and the range check are read back once here (pack time).
        self.tc_blob = None
        tc_ptr, s2max, tc_gamma = None, 0.0, 0.0
        nbytes = C.c_int64()
        if (kernel is not None and kernel.kind == _lib.DC_K_RQ and kernel.order == 2 and kernel.param > 0 and
                lib.dc_supports_tc_bytes(self.n, self.n_features, self.n_class, code, C.byref(nbytes)) == 0):
            blob = torch.empty(nbytes.value + 128, dtype=torch.uint8, device=device)
            ptr = (blob.data_ptr() + 127) // 128 * 128
            s2, valid = C.c_double(), C.c_int32()
            with torch.cuda.device(device):
                _lib.check(lib.dc_pack_supports_tc(s.data_ptr(), w.data_ptr(), self.n, self.n_features, C.byref(kernel), ptr,
                                                   _stream_ptr(device)), "dc_pack_supports_tc")
                torch.cuda.current_stream(device).synchronize()
                _lib.check(lib.dc_supports_tc_info(ptr, self.n, self.n_features, C.byref(s2), C.byref(valid)), "dc_supports_tc_info")
            if valid.value:
                self.tc_blob, tc_ptr, s2max, tc_gamma = blob, ptr, s2.value, float(kernel.param)
        # Optional low parts of the features (fk_forward_split): what float32 rounding dropped from FK(support) — the
        # tensor-core kernel's exact near-pair path adds them to its differences.
        self.table_lo = None
        if s_lo is not None and tc_ptr is not None and dtype == torch.float32:
            lo = s_lo.detach().reshape(self.n, -1).to(device=device, dtype=dtype).contiguous()
            if lo.shape != s.shape:
                raise ValueError(f"s_lo has shape {tuple(lo.shape)}, features have {tuple(s.shape)}")
            self.table_lo = torch.empty((self.n, row.value), dtype=dtype, device=device)
            with torch.cuda.device(device):
                _lib.check(lib.dc_pack_supports_lo(lo.data_ptr(), self.n, self.n_features, self.n_class,
                                                   self.table_lo.data_ptr(), _stream_ptr(device)), "dc_pack_supports_lo")
        self.desc = Supports(self.table.data_ptr(), self.n, self.n_features, self.n_class, f_pad.value, row.value, code, 0,
                             tc_ptr, s2max, tc_gamma, _ptr(self.table_lo))


def score_grad(fk: FkDesc, kernel: KernelDesc, sv: SupportSet, q: torch.Tensor, grad_mode: int = DC_GRAD_NONE,
               grad_out: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None
               ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """One launch of the fused hot path.  q: (B, D) on ``sv.device`` with ``sv.dtype``.

    Returns ``score`` (B, C) and, per ``grad_mode``, ``None`` | grad (B, D) | Jacobian (B, C, D).  ``out``: optional
    preallocated (B, C + D) [DC_GRAD_SUM] / (B, C) [DC_GRAD_NONE] buffer (any row stride) that receives the fused
    record [score | grad]; the returned tensors are then views into it.
    """
    lib = _lib.load()
    if q.device != sv.device or q.dtype != sv.dtype:
        raise ValueError(f"q must be {sv.dtype} on {sv.device}, got {q.dtype} on {q.device}")
    if q.ndim != 2 or q.shape[1] != fk.dof:
        raise ValueError(f"q must have shape (B, {fk.dof}), got {tuple(q.shape)}")
    if fk.n_features != sv.n_features:
        raise ValueError(f"feature map yields {fk.n_features} features, support set has {sv.n_features}")
    q = q.detach().contiguous()
    B, Cn, D = q.shape[0], sv.n_class, fk.dof
    grad = None
    if out is not None:
        if grad_mode == DC_GRAD_JAC:
            raise ValueError("out= is not supported with DC_GRAD_JAC")
        cols = Cn + (D if grad_mode == DC_GRAD_SUM else 0)
        if out.shape != (B, cols) or out.dtype != sv.dtype or out.device != sv.device or (B > 1 and out.stride(1) != 1):
            raise ValueError(f"out must be a ({B}, {cols}) {sv.dtype} tensor on {sv.device} with unit column stride")
        score = out[:, :Cn]
        grad = out[:, Cn:] if grad_mode == DC_GRAD_SUM else None
        score_ld = grad_ld = out.stride(0) if B > 1 else cols
    else:
        score = torch.empty((B, Cn), dtype=sv.dtype, device=sv.device)
        if grad_mode == DC_GRAD_SUM:
            grad = torch.empty((B, D), dtype=sv.dtype, device=sv.device)
        elif grad_mode == DC_GRAD_JAC:
            grad = torch.empty((B, Cn, D), dtype=sv.dtype, device=sv.device)
        score_ld = grad_ld = 0
    go = None
    if grad_out is not None:
        if grad_mode != DC_GRAD_SUM:
            raise ValueError("grad_out is only meaningful with DC_GRAD_SUM")
        go = grad_out.detach().to(device=sv.device, dtype=sv.dtype).reshape(B, Cn).contiguous()
    if B == 0:
        return score, grad
    with torch.cuda.device(sv.device):
        st = lib.dc_score_grad(C.byref(fk), C.byref(kernel), C.byref(sv.desc), q.data_ptr(), B, score.data_ptr(), score_ld,
                               _ptr(grad), grad_ld, _ptr(go), grad_mode, _stream_ptr(sv.device))
    _lib.check(st, "dc_score_grad")
    return score, grad


def kernel_matrix(kernel: KernelDesc, xa: torch.Tensor, xb: torch.Tensor) -> torch.Tensor:
    """K[i, j] = k(|xa_i - xb_j|^2) on flattened feature rows (training rows, fit_poly, jump-start block)."""
    lib = _lib.load()
    device = _require_cuda() if not xa.is_cuda else xa.device
    dtype = xa.dtype
    code = _dtype_code(dtype)
    a = xa.detach().reshape(xa.shape[0], -1).to(device=device, dtype=dtype).contiguous()
    b = xb.detach().reshape(xb.shape[0], -1).to(device=device, dtype=dtype).contiguous()
    if a.shape[1] != b.shape[1]:
        raise ValueError(f"feature mismatch: {a.shape[1]} vs {b.shape[1]}")
    out = torch.empty((a.shape[0], b.shape[0]), dtype=dtype, device=device)
    if out.numel():
        with torch.cuda.device(device):
            _lib.check(lib.dc_kernel_matrix(C.byref(kernel), a.data_ptr(), a.shape[0], b.data_ptr(), b.shape[0], a.shape[1],
                                            code, out.data_ptr(), _stream_ptr(device)), "dc_kernel_matrix")
    return out


def fk_forward(fk: FkDesc, q: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    q = q.detach().contiguous()
    out = torch.empty((q.shape[0], fk.n_features), dtype=q.dtype, device=q.device)
    if q.shape[0]:
        with torch.cuda.device(q.device):
            _lib.check(lib.dc_fk_forward(C.byref(fk), q.data_ptr(), q.shape[0], _dtype_code(q.dtype), out.data_ptr(),
                                         _stream_ptr(q.device)), "dc_fk_forward")
    return out


def fk_forward_split(fk: FkDesc, q: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """float32 features as (hi, lo): FK(q) = hi + lo to ~1e-9, hi == fk_forward(fk, q)."""
    lib = _lib.load()
    q = q.detach().to(torch.float32).contiguous()
    hi = torch.empty((q.shape[0], fk.n_features), dtype=torch.float32, device=q.device)
    lo = torch.empty_like(hi)
    if q.shape[0]:
        with torch.cuda.device(q.device):
            _lib.check(lib.dc_fk_forward_split(C.byref(fk), q.data_ptr(), q.shape[0], hi.data_ptr(), lo.data_ptr(),
                                               _stream_ptr(q.device)), "dc_fk_forward_split")
    return hi, lo


def fk_vjp(fk: FkDesc, q: torch.Tensor, g_x: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    q = q.detach().contiguous()
    g_x = g_x.detach().to(dtype=q.dtype, device=q.device).reshape(q.shape[0], fk.n_features).contiguous()
    out = torch.empty_like(q)
    if q.shape[0]:
        with torch.cuda.device(q.device):
            _lib.check(lib.dc_fk_vjp(C.byref(fk), q.data_ptr(), q.shape[0], _dtype_code(q.dtype), g_x.data_ptr(),
                                     out.data_ptr(), _stream_ptr(q.device)), "dc_fk_vjp")
    return out


def fk_tree_frames(fk: FkDesc, q: torch.Tensor) -> torch.Tensor:
    """(B, n_nodes, 12): row-major [R | t] of every body of a DC_FK_JOINT_TREE map; computed on the GPU, returned where
    ``q`` lives."""
    lib = _lib.load()
    dev = q.device if q.is_cuda else _require_cuda()
    dtype = q.dtype if q.dtype in (torch.float32, torch.float64) else torch.float32
    qd = q.detach().to(device=dev, dtype=dtype).reshape(-1, fk.dof).contiguous()
    out = torch.empty((qd.shape[0], fk.n_nodes, 12), dtype=dtype, device=dev)
    if qd.shape[0]:
        with torch.cuda.device(dev):
            _lib.check(lib.dc_fk_tree_frames(C.byref(fk), qd.data_ptr(), qd.shape[0], _dtype_code(dtype), out.data_ptr(),
                                             _stream_ptr(dev)), "dc_fk_tree_frames")
    return out.to(q.device)


# --------------------------------------------------------------------------------------------------------
# autograd
# --------------------------------------------------------------------------------------------------------


def _is_batched(t) -> bool:
    """True inside a vmapped backward (``jacobian(vectorize=True)``, ``is_grads_batched``): such tensors have no storage a
    CUDA launch could read."""
    try:  # both vmap implementations mark their tensors with a dispatch key (Batched / FuncTorchBatched)
        return "Batched" in str(torch._C._dispatch_keys(t))
    except Exception:
        return False


class _ScoreFunction(torch.autograd.Function):
    """score = evaluator(q) with an analytic backward.

    ``evaluator(q, mode, grad_out) -> (score (B, C), grad)`` runs the fused CUDA kernel (modes of ``dc_score_grad``).

    * One class: forward is ONE launch that also returns the Jacobian ``(B, 1, D)``; backward is a multiply.
    * Several classes: forward returns the scores only.  backward launches the kernel once more with the upstream gradient
      (``DC_GRAD_SUM``: ``(B, D)`` out) — not the ``(B, C, D)`` Jacobian, C times the output the optimisers need
      (diffco/optim.py:101,734 differentiate a sum).  Only when autograd batches the upstream gradient
      (``torch.autograd.functional.jacobian(vectorize=True)``, diffco/optim.py:211-216) is the full Jacobian evaluated —
      on the unbatched ``q`` — and contracted with pure torch ops (mul + sum have batching rules, a CUDA launch does not).

    The gradient is a constant w.r.t. autograd: second derivatives (optim.py:380-391) are not provided through autograd
    (``optim.trustconstr_traj_optimize`` takes finite differences of this first derivative) and fail loudly instead of
    silently returning zeros.
    """

    @staticmethod
    def forward(q, evaluator):
        want = q.requires_grad or torch.is_grad_enabled()
        if want and getattr(evaluator, "n_class", 1) == 1:
            return evaluator(q, _lib.DC_GRAD_JAC)
        score, _ = evaluator(q, _lib.DC_GRAD_NONE)
        return score, None

    @staticmethod
    def setup_context(ctx, inputs, output):
        q, evaluator = inputs
        _, jac = output
        ctx.evaluator = evaluator
        if jac is not None:
            ctx.save_for_backward(jac)
            ctx.mark_non_differentiable(jac)
        else:
            ctx.save_for_backward(q)
        ctx.has_jac = jac is not None
        ctx.set_materialize_grads(False)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_score, _grad_jac):
        if grad_score is None:
            return None, None
        if ctx.has_jac:
            (jac,) = ctx.saved_tensors
        else:
            (q,) = ctx.saved_tensors
            if not _is_batched(grad_score):
                _, grad = ctx.evaluator(q, _lib.DC_GRAD_SUM, grad_score)
                return grad.to(q.dtype), None
            _, jac = ctx.evaluator(q, _lib.DC_GRAD_JAC)
        # mul + sum (not einsum / bmm): both have batching rules, which the vmapped backward needs
        if jac.shape[-2] == 1:  # one class: (B, 1) * (B, D), a single elementwise kernel
            return grad_score.to(jac.dtype) * jac[..., 0, :], None
        return (grad_score.to(jac.dtype).unsqueeze(-1) * jac).sum(-2), None


def differentiable_score(q: torch.Tensor, evaluator) -> torch.Tensor:
    """Run ``evaluator`` on q (any device / dtype handled by the evaluator) with autograd support."""
    if q.requires_grad and torch.is_grad_enabled():
        score, _ = _ScoreFunction.apply(q, evaluator)
        return score
    score, _ = evaluator(q, _lib.DC_GRAD_NONE)
    return score


class _FkFunction(torch.autograd.Function):
    """fkine with the J^T product from dc_fk_vjp (first derivatives only)."""

    @staticmethod
    def forward(q, runner):
        return runner.forward(q)

    @staticmethod
    def setup_context(ctx, inputs, output):
        q, runner = inputs
        ctx.runner = runner
        ctx.save_for_backward(q)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_x):
        (q,) = ctx.saved_tensors
        return ctx.runner.vjp(q, grad_x), None


# --------------------------------------------------------------------------------------------------------
# host buffers in / host buffers out
# --------------------------------------------------------------------------------------------------------


class HostPipeline:
    """Owner of a native ``dc_host_pipeline`` (include/diffco_b200.h): a few CUDA streams + device staging buffers through
    which ``dc_score_grad_host`` pipelines H2D copy -> fused kernel -> D2H copy chunk by chunk.  The streams fork from and
    join back into the caller's current stream, so events recorded there bracket all of the work and the call never
    synchronises with the host."""

    def __init__(self, device: torch.device, chunk_rows: int = 16384, n_slots: int = 3):
        self.device = device
        self.chunk_rows = chunk_rows
        self._lib = _lib.load()
        handle = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(self._lib.dc_host_pipeline_create(C.byref(handle), chunk_rows, n_slots), "dc_host_pipeline_create")
        self._handle = handle

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h is not None and h.value:
            try:
                self._lib.dc_host_pipeline_destroy(h)
            except Exception:
                pass

    def score_grad(self, fk: FkDesc, kernel: KernelDesc, sv: SupportSet, q_host: torch.Tensor, out_host: torch.Tensor,
                   grad_mode: int = DC_GRAD_SUM) -> torch.Tensor:
        """q_host (B, D), out_host (B, C [+ D]) contiguous CPU tensors of the model dtype (pinned for overlap)."""
        if q_host.is_cuda or out_host.is_cuda:
            raise ValueError("score_grad_host takes CPU tensors; use score_grad for device-resident batches")
        B = q_host.shape[0]
        rec = sv.n_class + (fk.dof if grad_mode == DC_GRAD_SUM else 0)
        if q_host.dtype != sv.dtype or out_host.dtype != sv.dtype or not q_host.is_contiguous() or not out_host.is_contiguous():
            raise ValueError(f"host buffers must be contiguous {sv.dtype} tensors")
        if q_host.shape != (B, fk.dof) or out_host.shape != (B, rec):
            raise ValueError(f"expected q_host ({B}, {fk.dof}) and out_host ({B}, {rec})")
        if B:
            with torch.cuda.device(self.device):
                st = self._lib.dc_score_grad_host(self._handle, C.byref(fk), C.byref(kernel), C.byref(sv.desc), q_host.data_ptr(),
                                                  B, out_host.data_ptr(), grad_mode, _stream_ptr(self.device))
            _lib.check(st, "dc_score_grad_host")
        return out_host
