"""Radial kernels — host-side mirror of the reference's ``diffco/kernel.py``.

Each class keeps the reference constructor and call signature (``kernel(xs, x_primes) -> (B, N)`` including the
reference's squeeze quirks) and additionally exposes ``desc``: the ``dc_kernel_desc`` POD that the fused CUDA
kernels evaluate per pair in registers.  Calling a kernel object computes the kernel matrix with
``dc_kernel_matrix`` on the GPU (used by ``train`` / ``fit_poly``); inside ``DiffCo.score`` the kernel is never
materialised at all.
"""
from __future__ import annotations

import torch

from . import _lib, functional
from ._lib import KernelDesc


class KernelFunc:
    """kernel.py:4-9."""

    desc: KernelDesc = None
    squeeze_single_row = False  # RQKernel / MultiQuadratic squeeze a (1, N) result to (N,) (kernel.py:26-27,55-56)

    def __call__(self, xs, x_primes):
        raise NotImplementedError("You need to define your own __call__ function.")

    def _matrix(self, xs, x_primes):
        """K (B, N) on the GPU, returned on x_primes' device.  Shapes follow kernel.py:18-21: leading dims are
        prepended to xs until ranks match, everything after dim 0 is flattened."""
        while xs.ndim < x_primes.ndim:
            xs = xs.unsqueeze(0)
        dev = x_primes.device
        dtype = x_primes.dtype if x_primes.dtype in (torch.float32, torch.float64) else torch.float32
        cuda = functional._require_cuda() if not x_primes.is_cuda else dev
        a = xs.detach().reshape(xs.shape[0], -1).to(device=cuda, dtype=dtype)
        b = x_primes.detach().reshape(x_primes.shape[0], -1).to(device=cuda, dtype=dtype)
        k = functional.kernel_matrix(self.desc, a, b).to(dev)
        if self.squeeze_single_row and k.shape[0] == 1:
            k = k.squeeze(0)
        return k


class RQKernel(KernelFunc):
    """kernel.py:12-29 — k = (1 + gamma/p |x-s|^2)^-p."""

    squeeze_single_row = True

    def __init__(self, gamma: float, p: int = 2):
        if int(p) != p or p < 1:
            raise ValueError("RQKernel: p must be a positive integer")
        self.gamma = gamma
        self.p = int(p)
        self.desc = KernelDesc(_lib.DC_K_RQ, self.p, float(gamma))

    def __call__(self, xs, x_primes):
        return self._matrix(xs, x_primes)


class MultiQuadratic(KernelFunc):
    """kernel.py:45-57 — k = sqrt(|x-s|^2/eps^2 + 1)."""

    squeeze_single_row = True

    def __init__(self, epsilon):
        self.epsilon = epsilon
        self.desc = KernelDesc(_lib.DC_K_MULTIQUADRIC, 0, float(epsilon))

    def __call__(self, xs, x_primes):
        if xs.ndim == 1:
            xs = xs[None, :]
        return self._matrix(xs, x_primes)


class Polyharmonic(KernelFunc):
    """kernel.py:59-79 — k = r^k/eps (k odd), r^k log r/eps with 0 at r=0 (k even); never squeezed."""

    def __init__(self, k, epsilon):
        if int(k) != k or k < 1:
            raise ValueError("Polyharmonic: k must be a positive integer")
        self.k = int(k)
        self.epsilon = epsilon
        self.desc = KernelDesc(_lib.DC_K_POLYHARMONIC, self.k, float(epsilon))

    def __call__(self, xs, x_primes):
        return self._matrix(xs, x_primes)


class FKKernel(KernelFunc):
    """kernel.py:131-143.  Deprecated in the reference (its constructor raises); kept because every legacy script
    builds its perceptron as ``DiffCo(obstacles, kernel_func=FKKernel(fkine, RQKernel(gamma)))``.  Here it is only
    a carrier: ``DiffCo`` / ``MultiDiffCo`` unpack it into (transform, radial kernel) and fuse both."""

    def __init__(self, fkine, rq_kernel):
        self.fkine = fkine
        self.rq_kernel = rq_kernel
        self.desc = rq_kernel.desc
        self.squeeze_single_row = rq_kernel.squeeze_single_row

    def __call__(self, xs, x_primes=None, x_primes_controls=None):
        if xs.ndim == 1:
            xs = xs[None, :]
        xs_controls = self.fkine(xs).reshape(len(xs), -1)
        if x_primes_controls is None:
            x_primes_controls = self.fkine(x_primes).reshape(len(x_primes), -1)
        return self.rq_kernel(xs_controls, x_primes_controls)


def _robot_of(fkine):
    owner = getattr(fkine, "__self__", None)
    if owner is None or getattr(owner, "fk_desc", None) is None:
        raise TypeError("fkine must be the bound .fkine of a diffco_b200.model robot")
    return owner


class _TemporalRQ(KernelFunc):
    """The radial part of TemporalFKKernel on features [x | t]: RQ(gamma, p)(|dx|^2) * RQ(gamma_t, p_t)(dt^2)^alpha
    (DC_K_RQ_TEMPORAL, evaluated per pair in registers like every other kernel)."""

    squeeze_single_row = True  # both factors are RQKernel results (kernel.py:26-27)

    def __init__(self, rq, t_rq, alpha):
        if not isinstance(rq, RQKernel) or not isinstance(t_rq, RQKernel):
            raise TypeError("TemporalFKKernel: both kernels must be RQKernel")
        self.desc = KernelDesc(_lib.DC_K_RQ_TEMPORAL, rq.p, float(rq.gamma), float(t_rq.gamma), float(alpha), t_rq.p, 0)

    def __call__(self, xs, x_primes):
        return self._matrix(xs, x_primes)


class TemporalFKKernel(KernelFunc):
    """kernel.py:175-202 — k((q1,t1),(q2,t2)) = rq(FK(q1), FK(q2)) * t_rq(t1, t2)^alpha; time is the last column.
    ``DiffCo`` unpacks it into (``map``: [q | t] -> [FK(q) | t], ``radial``) and fuses both into the score kernel."""

    def __init__(self, fkine, rqkernel, t_rqkernel, alpha=0.5):
        from .model import ComposedMap

        self.fkine = fkine
        self.rqkernel = rqkernel
        self.t_rqkernel = t_rqkernel
        self.alpha = alpha
        self.map = ComposedMap(_robot_of(fkine), 1, time_last=True)
        self.radial = _TemporalRQ(rqkernel, t_rqkernel, alpha)
        self.desc = self.radial.desc

    def __call__(self, xs, x_primes):
        if xs.ndim == 1:
            xs = xs[None, :]
        return self.radial(self.map.fkine(xs).reshape(len(xs), -1), self.map.fkine(x_primes).reshape(len(x_primes), -1))


class LineFKKernel(KernelFunc):
    """kernel.py:188-202 — rq on the concatenated features of a segment's two end configurations ([q_a | q_b] rows).
    ``DiffCo`` unpacks it into (``map``: FK on both halves, ``radial`` = the RQ kernel)."""

    def __init__(self, fkine, rq_kernel):
        from .model import ComposedMap

        self.fkine = fkine
        self.rq_kernel = rq_kernel
        self.map = ComposedMap(_robot_of(fkine), 2)
        self.radial = rq_kernel
        self.desc = rq_kernel.desc

    def __call__(self, xs, x_primes):
        if xs.ndim == 1:
            xs = xs[None, :]
        if x_primes.ndim == 1:
            x_primes = x_primes[None, :]
        if xs.shape[1] != x_primes.shape[1] or xs.shape[1] % 2:
            raise AssertionError("LineFKKernel: rows are [q_a | q_b]")
        return self.rq_kernel(self.map.fkine(xs).reshape(len(xs), -1), self.map.fkine(x_primes).reshape(len(x_primes), -1))


class LineKernel(KernelFunc):
    """kernel.py:175-186 — mean of a point kernel on the two halves of [q_a | q_b] rows.  A sum of two radial kernels
    is not radial, so it has no fused descriptor: usable as a kernel matrix (two dc_kernel_matrix launches), not as
    ``DiffCo``'s score kernel."""

    def __init__(self, point_kernel):
        self.point_kernel = point_kernel

    def __call__(self, xs, x_primes):
        if xs.ndim == 1:
            xs = xs[None, :]
        if x_primes.ndim == 1:
            x_primes = x_primes[None, :]
        if xs.shape[1] != x_primes.shape[1] or xs.shape[1] % 2:
            raise AssertionError("LineKernel: rows are [q_a | q_b]")
        dof = xs.shape[1] // 2
        pk = self.point_kernel
        return (pk._matrix(xs[:, :dof], x_primes[:, :dof]) + pk._matrix(xs[:, dof:], x_primes[:, dof:])) / 2
