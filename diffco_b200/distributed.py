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

import ctypes as C
import os
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib, functional
from ._lib import DC_GRAD_SUM


def bind_host_thread_to_gpu(device_index: int) -> Optional[list]:
    """Pin the calling thread to the CPUs that are NUMA-local to GPU ``device_index`` (physical index, i.e. after
    CUDA_VISIBLE_DEVICES), so that pinned host buffers allocated afterwards land on that GPU's memory node (first touch).
    The zero-copy host path (``score_and_grad_host``) moves every query and record over PCIe inside the kernel; with eight
    ranks on a two-socket box, buffers on the wrong socket halve its throughput.  Returns the CPU list, or None when the
    topology cannot be read (then nothing is changed)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device_index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        with open(f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = []
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.extend(range(int(lo), int(hi or lo) + 1))
        cpus = sorted(set(cpus) & os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


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


class _DevicePtr:
    """Minimal __cuda_array_interface__ holder so torch can view memory owned by libdiffco_b200 (dc_peer_alloc)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False), "version": 2}


class PeerExchange:
    """The gathered (G*b, C+D) record buffer of every rank of one box, mapped into every process (CUDA IPC over NVLink).

    ``dc_score_grad_bcast`` stores each tile's records into all G buffers from the kernel's epilogue, so the all-gather
    is fused into the scoring kernel; ``dc_peer_barrier`` (one flag round trip) then publishes the step.  Two buffers
    alternate, so a rank that races ahead into step k+1 never overwrites records a slower rank still reads from step k
    (it cannot reach step k+2 before everyone has passed the barrier of step k+1).
    """

    FLAG_BYTES = 256

    def __init__(self, rows: int, width: int, dtype: torch.dtype, device: torch.device, group):
        assert dtype == torch.float32
        self.lib = _lib.load()
        self.group, self.world, self.rank = group, dist.get_world_size(group), dist.get_rank(group)
        if self.world > _lib.DC_MAX_PEERS:
            raise RuntimeError(f"at most {_lib.DC_MAX_PEERS} ranks")
        self.rows, self.width, self.device = rows, width, device
        # each of the two record buffers starts on a 128-byte boundary (the kernel's bulk stores need 16-byte aligned
        # destinations whatever rows * width is)
        self.buf_bytes = -(-(rows * width * 4) // 128) * 128
        self.nbytes = self.FLAG_BYTES + 2 * self.buf_bytes
        self._own, self._opened = C.c_void_p(), []
        # Every rank walks through the SAME sequence of collectives whatever fails locally (allocation, IPC mapping):
        # failures are folded into one consensus at the end, which doubles as the "everyone has mapped every buffer"
        # barrier; then all ranks raise together and the caller falls back to NCCL on all of them.
        ok, mine = True, None
        try:
            handle = _lib.PeerHandle()
            with torch.cuda.device(device):
                _lib.check(self.lib.dc_peer_alloc(self.nbytes, C.byref(self._own), C.byref(handle)), "dc_peer_alloc")
            mine = bytes(handle.bytes)
        except Exception:
            ok = False
        handles = [None] * self.world
        dist.all_gather_object(handles, mine, group=group)
        bases = []
        if ok and all(h is not None for h in handles):
            try:
                for r, hb in enumerate(handles):
                    if r == self.rank:
                        bases.append(self._own.value)
                        continue
                    h = _lib.PeerHandle()
                    C.memmove(h.bytes, hb, 64)
                    ptr = C.c_void_p()
                    with torch.cuda.device(device):
                        _lib.check(self.lib.dc_peer_open(C.byref(h), C.byref(ptr)), "dc_peer_open")
                    self._opened.append(ptr)
                    bases.append(ptr.value)
            except Exception:
                ok = False
        else:
            ok = False
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device if dist.get_backend(group) == "nccl" else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            self.close()
            raise RuntimeError("peer memory (CUDA IPC) is not available on every rank")
        self.flags = _lib.PeerTable()
        self.outs = [_lib.PeerTable(), _lib.PeerTable()]
        for r, base in enumerate(bases):
            self.flags.ptr[r] = base
            for k in range(2):
                self.outs[k].ptr[r] = base + self.FLAG_BYTES + k * self.buf_bytes
        self.views = [torch.as_tensor(_DevicePtr(self._own.value + self.FLAG_BYTES + k * self.buf_bytes, (rows, width), "<f4"),
                                      device=device) for k in range(2)]
        self.epoch = 0   # barriers so far
        self.steps = 0   # scoring steps so far (selects the buffer)
        # DIFFCO_B200_PEER_SYNC=1: the step barrier runs in the tail of the scoring kernel (dc_score_grad_bcast_sync, one launch
        # per step) instead of a second launch.  Measured at 2 GPUs: 0.1111 vs 0.1101 ms per step — the launch gap is not what
        # the barrier costs (profiles/r03m_barrier_variants.txt) — so two launches stay the default.
        self.fold_barrier = os.environ.get("DIFFCO_B200_PEER_SYNC", "0") == "1"

    def release(self):
        """Collective: every rank stops using the mapped buffers, THEN they are unmapped and freed (a rank must not free
        memory a slower peer's kernel may still be storing into)."""
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        self.close()

    def close(self):
        for p in getattr(self, "_opened", []):
            self.lib.dc_peer_close(p)
        self._opened = []
        if getattr(self, "_own", None) is not None and self._own.value:
            self.lib.dc_peer_free(self._own)
            self._own = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def score_and_grad(self, fk, kdesc, sv, q_shard: torch.Tensor, mirror: Optional[torch.Tensor] = None
                       ) -> Optional[torch.Tensor]:
        """Scores this rank's shard into every rank's buffer and publishes it.  ``q_shard`` lives on the device or in pinned
        host memory (read zero-copy); ``mirror`` (optional, (b, C+D), device or pinned host) receives a copy of this rank's
        records.  Returns the (G*b, C+D) view holding the global result, or None when the call is not one the tensor-core
        kernel takes (the answer depends on shapes and options only, so every rank gets the same one and the caller falls
        back to NCCL on all of them)."""
        b = q_shard.shape[0]
        k = self.steps & 1
        stream = functional._stream_ptr(self.device)
        mptr = None if mirror is None else mirror.data_ptr()
        with torch.cuda.device(self.device):
            if self.fold_barrier:
                # ONE launch: the last CTA of the grid runs the flag exchange (dc_score_grad_bcast_sync)
                st = self.lib.dc_score_grad_bcast_sync(C.byref(fk), C.byref(kdesc), C.byref(sv.desc), q_shard.data_ptr(), b,
                                                       C.byref(self.outs[k]), self.world, self.rank * b, DC_GRAD_SUM, mptr,
                                                       C.byref(self.flags), self.rank, self.epoch + 1, stream)
            else:
                st = self.lib.dc_score_grad_bcast(C.byref(fk), C.byref(kdesc), C.byref(sv.desc), q_shard.data_ptr(), b,
                                                  C.byref(self.outs[k]), self.world, self.rank * b, DC_GRAD_SUM, mptr, stream)
            if st == -2:  # DC_ERR_UNSUPPORTED
                return None
            _lib.check(st, "dc_score_grad_bcast")
        self.steps += 1
        if self.fold_barrier:
            self.epoch += 1
        else:
            self.barrier()
        return self.views[k]

    def barrier(self):
        """Device-side barrier of all ranks on the current stream (one flag round trip, no host involvement)."""
        self.epoch += 1
        with torch.cuda.device(self.device):
            _lib.check(self.lib.dc_peer_barrier(C.byref(self.flags), self.rank, self.world, self.epoch,
                                                functional._stream_ptr(self.device)), "dc_peer_barrier")


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
        self._peer = None          # PeerExchange for the current shard size (fused all-gather), or False: unavailable
        self._peer_rows = None
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
        rec = self.gathered_records(q_shard)
        return rec[:, :self.n_class], rec[:, self.n_class:]

    def gathered_records(self, q_shard: torch.Tensor, mirror: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The gathered (G*b, C+D) [score | grad] records of all ranks' shards (a view of the buffer the collective filled).
        With the fused all-gather ``q_shard`` may be pinned host memory and ``mirror`` receives this rank's records."""
        b = q_shard.shape[0]
        if self.world > 1 and self._local_fn is None and self.dtype == torch.float32 and (q_shard.is_cuda or q_shard.is_pinned()):
            out = self._fused_all_gather(q_shard, b, mirror)
            if out is not None:
                return out
        if not q_shard.is_cuda and self.device.type == "cuda":
            if self._q_stage is None or self._q_stage.shape[0] < b:
                self._q_stage = torch.empty((b, self.dof), dtype=self.dtype, device=self.device)
            q_dev = self._q_stage[:b]
            q_dev.copy_(q_shard, non_blocking=True)
            q_shard = q_dev
        buf = self._buffer(self.world * b)
        self._local(q_shard, buf[self.rank * b:(self.rank + 1) * b])
        if self.world > 1:
            all_gather_rows(buf, b, self.rank, self.group)
        if mirror is not None:
            mirror.copy_(buf[self.rank * b:(self.rank + 1) * b], non_blocking=True)
        return buf

    def align(self) -> None:
        """Stream-ordered rendezvous of all ranks (benchmarks: line the ranks up after untimed work)."""
        if self.world == 1:
            return
        if self._peer:
            self._peer.barrier()
        else:
            t = torch.zeros(1, device=self.device)
            dist.all_reduce(t, group=self.group)

    def _fused_all_gather(self, q_shard: torch.Tensor, b: int, mirror: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
        """Tensor-core kernel with the all-gather fused into its epilogue (peer stores over NVLink + one flag barrier);
        None when unavailable (DIFFCO_B200_PEER=0, IPC mapping failed, or the call is not a tensor-core one)."""
        if self._peer is False or os.environ.get("DIFFCO_B200_PEER", "1") == "0":
            return None
        if self._peer is None or self._peer_rows != b:
            if self._peer:
                self._peer.release()  # collective, like the construction below: every rank sees the same shard size change
            try:  # PeerExchange raises on ALL ranks or on none (its constructor ends with a consensus)
                self._peer = PeerExchange(self.world * b, self.record_width, self.dtype, self.device, self.group)
                self._peer_rows = b
            except RuntimeError:
                self._peer = False
                return None
        if mirror is not None and not (mirror.is_cuda or mirror.is_pinned()):
            return None
        return self._peer.score_and_grad(self._fk, self._kfun.desc, self._sv, q_shard.detach().contiguous(), mirror)

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
        if self.world > 1:
            # one launch with pinned buffers (the kernel reads q over PCIe, stores this rank's records to out_host AND to
            # every peer's gathered buffer); staged copies around the collective otherwise
            self.gathered_records(q_host, mirror=out_host)
            return out_host
        if self._pipe is None:
            self._pipe = functional.HostPipeline(self.device)
        return self._pipe.score_grad(self._fk, self._kfun.desc, self._sv, q_host, out_host, DC_GRAD_SUM)
