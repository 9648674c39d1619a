"""Multi-GPU scoring: the batch dimension is sharded across the ranks of one box, supports are replicated.

The reference has no distributed code; queries are independent and the support table is tiny
(<= 0.4 MB), so the natural partition is contiguous rows of the batch: rank r owns rows [r*ceil(B/G), (r+1)*ceil(B/G)).
Each rank's fused kernel writes its [score | grad] records straight into its slice of the gathered (G*b, C+D) buffer
(``dc_score_grad`` with score_ld = grad_ld = C+D) and ONE all-gather (NCCL over NVLink / NVSwitch, in place) makes the
whole batch visible on every rank — no staging copies, no second collective.

The host-side logic (partitioning, padding, buffer layout, trimming) is independent of the device and is covered
by world_size-2 gloo tests with a stub local scorer; the product path always computes with the CUDA kernel.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

from . import functional
from ._lib import DC_GRAD_SUM


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int, int]:
    """Contiguous balanced partition: (lo, hi, rows_per_rank) with rows_per_rank = ceil(total / world); the last
    ranks may own fewer (or zero) real rows and are padded up to rows_per_rank for the equal-size all-gather."""
    per = -(-total // world) if total > 0 else 0
    lo = min(total, rank * per)
    hi = min(total, lo + per)
    return lo, hi, per


def all_gather_rows(buf: torch.Tensor, per: int, rank: int, group=None) -> None:
    """In-place all-gather of equal row blocks: rank r's rows [r*per, (r+1)*per) of ``buf`` are sent to every rank."""
    if per == 0:
        return
    backend = dist.get_backend(group)
    mine = buf[rank * per:(rank + 1) * per]
    if backend == "nccl":
        dist.all_gather_into_tensor(buf, mine, group=group)
    else:  # gloo (CPU tests): list form, same result
        world = dist.get_world_size(group)
        views = [buf[r * per:(r + 1) * per] for r in range(world)]
        dist.all_gather(views, mine.clone(), group=group)


class ShardedScorer:
    """score + gradient of a (replicated) perceptron over a batch sharded across the ranks of ``group``.

    ``checker``: a diffco_b200 DiffCo / MultiDiffCo; ``weights``: 'gains' (score) or 'rbf' (poly_score / rbf_score).
    ``local_fn(q, out)`` can be injected for host-logic tests; by default it is the fused CUDA launch.
    """

    def __init__(self, checker, weights: str = "gains", group=None,
                 local_fn: Optional[Callable[[torch.Tensor, torch.Tensor], None]] = None):
        self.group = group
        self.distributed = group is not None and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.distributed else 1
        self.rank = dist.get_rank(group) if self.distributed else 0
        self._checker = checker
        self._weights = weights
        self._local_fn = local_fn
        self._buffers = {}
        self._pipe = None
        self._q_stage = None
        if local_fn is None:
            self._sv, self._kfun = checker._select(weights)
            self._fk = checker._fk_for(self._sv)
            self.n_class, self.dof = self._sv.n_class, self._fk.dof
            self.dtype, self.device = self._sv.dtype, self._sv.device
        else:
            self.n_class, self.dof = checker.n_class, checker.dof
            self.dtype, self.device = checker.dtype, checker.device

    @property
    def record_width(self) -> int:
        return self.n_class + self.dof

    # ---------------------------------------------------------------- single-rank pieces
    def _local(self, q: torch.Tensor, out: torch.Tensor) -> None:
        if self._local_fn is not None:
            self._local_fn(q, out)
        else:
            functional.score_grad(self._fk, self._kfun.desc, self._sv, q, DC_GRAD_SUM, out=out)

    def _buffer(self, rows: int) -> torch.Tensor:
        buf = self._buffers.get(rows)
        if buf is None:
            buf = torch.empty((rows, self.record_width), dtype=self.dtype, device=self.device)
            self._buffers = {rows: buf}  # keep one: the optimisers call with a fixed batch
        return buf

    def local_score_and_grad(self, q: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """This rank's rows only (no collective): (score (b, C), grad (b, D)) as views of one (b, C+D) record buffer."""
        out = self._buffer(q.shape[0])
        self._local(q, out)
        return out[:, :self.n_class], out[:, self.n_class:]

    # ---------------------------------------------------------------- sharded entry points
    def score_and_grad(self, q_shard: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Every rank passes ITS shard (same row count b on every rank); returns the gathered global result
        (score (G*b, C), grad (G*b, D)) — rank r's rows are [r*b, (r+1)*b)."""
        b = q_shard.shape[0]
        buf = self._buffer(self.world * b)
        self._local(q_shard, buf[self.rank * b:(self.rank + 1) * b])
        if self.world > 1:
            all_gather_rows(buf, b, self.rank, self.group)
        return buf[:, :self.n_class], buf[:, self.n_class:]

    def score_and_grad_global(self, q_global: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Every rank holds the same global batch (B, D); rank r evaluates rows shard_bounds(B, G, r) and the result for
        all B rows is returned on every rank (ragged B is padded for the collective and trimmed afterwards)."""
        total = q_global.shape[0]
        lo, hi, per = shard_bounds(total, self.world, self.rank)
        buf = self._buffer(self.world * per)
        if hi > lo:
            self._local(q_global[lo:hi].contiguous(), buf[self.rank * per:self.rank * per + (hi - lo)])
        if hi - lo < per:
            buf[self.rank * per + (hi - lo):(self.rank + 1) * per].zero_()
        if self.world > 1:
            all_gather_rows(buf, per, self.rank, self.group)
        return buf[:total, :self.n_class], buf[:total, self.n_class:]

    # ---------------------------------------------------------------- host buffers in, host buffers out
    def score_and_grad_host(self, q_host: torch.Tensor, out_host: torch.Tensor) -> torch.Tensor:
        """End-to-end call for host-resident batches: ``q_host`` (b, D) and ``out_host`` (b, C+D) are (preferably pinned)
        CPU tensors.  Single rank: ``dc_score_grad_host`` — chunks flow H2D -> fused kernel -> D2H on the pipeline's
        streams so the PCIe copies overlap the compute.  With world > 1 the device-side result is also all-gathered (every
        rank keeps the global batch on its GPU) before this rank's rows are copied back.
        Returns ``out_host`` once the work is enqueued; the caller synchronises the current stream."""
        if self._local_fn is not None:
            raise RuntimeError("score_and_grad_host needs the CUDA scorer")
        b = q_host.shape[0]
        if self.world > 1:
            if self._q_stage is None or self._q_stage.shape[0] < b:
                self._q_stage = torch.empty((b, self.dof), dtype=self.dtype, device=self.device)
            qd = self._q_stage[:b]
            qd.copy_(q_host, non_blocking=True)
            self.score_and_grad(qd)
            buf = self._buffer(self.world * b)
            out_host.copy_(buf[self.rank * b:(self.rank + 1) * b], non_blocking=True)
            return out_host
        if self._pipe is None:
            self._pipe = functional.HostPipeline(self.device)
        return self._pipe.score_grad(self._fk, self._kfun.desc, self._sv, q_host, out_host, DC_GRAD_SUM)
