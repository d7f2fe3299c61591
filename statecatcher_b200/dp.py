"""Data parallelism by independent audio streams.

The reference has no distributed code at all; the only parallel axis the path offers is the
batch of streams: nothing inside LucyRNN (row-wise in b, lucyrnn.py:44-70) or CTC
(per-utterance lattice) mixes streams.  So each rank owns a contiguous block of streams and
the carried ``(h, s)`` of exactly those streams for the whole recording — state never crosses
ranks — and the single collective of a training step is a gradient all-reduce.

``StreamDataParallel`` keeps the gradients' slices of one persistent flat buffer per bucket (parameters in
reverse registration order = the order the backward pass finishes them: output_proj first, layer 0 last) and
exchanges them with an NCCL all-reduce — by default one collective from the end-of-backward callback, optionally
(``overlap=True``) one per bucket from a post-accumulate-grad hook the moment the bucket's last gradient lands.  Averaging by world size reproduces
``reduction='mean'`` over the GLOBAL batch when shards are equal (mean of rank means); for
unequal shards multiply the local loss by ``shard_loss_scale`` first.
"""
from __future__ import annotations

from typing import List, Optional

import ctypes
import os

import torch
import torch.distributed as dist
import torch.nn as nn


def partition_streams(n_streams: int, world_size: int, rank: int) -> range:
    """Contiguous, near-equal block of stream ids owned by ``rank`` (cfg3: 512 -> 64 each)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(n_streams, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def shard_loss_scale(n_local: int, n_global: int, world_size: int) -> float:
    """Factor for a rank's ``reduction='mean'`` loss when shards are NOT equal (n_streams not a
    multiple of the world size, or a rank's streams ran out): the wrapper averages rank
    gradients, i.e. computes ``(1/W) * sum_r mean_r``; scaling rank r's loss by
    ``n_local * W / n_global`` turns that into the mean over the global batch.  1.0 for equal shards."""
    if n_global <= 0 or world_size <= 0 or not (0 <= n_local <= n_global):
        raise ValueError("shard_loss_scale: need 0 <= n_local <= n_global, n_global > 0, world_size > 0")
    return n_local * world_size / n_global


class _Bucket:
    def __init__(self, params: List[nn.Parameter]):
        self.params = params
        self.pending = len(params)
        self.work = None
        self.flat: Optional[torch.Tensor] = None       # persistent communication buffer (allocated on first use)
        self.offsets: List[int] = []
        self.avg_in_collective = False
        self.live: List[nn.Parameter] = []             # params whose gradient went into the collective in flight


def _aligned_offsets(params, align=8):
    """Element offsets of the parameters' slices in a flat buffer, each slice starting on a 16-byte boundary
    (for 2-byte elements too)."""
    offs, cur = [], 0
    for p in params:
        offs.append(cur)
        cur += (p.numel() + align - 1) // align * align
    return offs, cur


class StreamDataParallel(nn.Module):
    """Gradient all-reduce over the ranks of a stream-sharded job.

    ``overlap=False`` (default since r02; env ``SC_DP_OVERLAP=1`` / ``overlap=True`` restores the hook-launched
    variant): every bucket is exchanged from the end-of-backward callback.  Measured on 8 x B200 at configs[2]
    (profiles/r01_bench_8gpu_variants.txt, r02_dp_variants.txt): the backward is a chain of persistent one-CTA-per-SM
    tcgen05 GEMMs and HBM-bound scans, an NCCL kernel launched under it either waits for SMs or takes bandwidth from
    the scan of the layer below (scan backward 5.18 -> 6.57 ms per step), and buys nothing — the exchange itself is
    ~0.3 ms of a 35 ms step.
    ``grad_dtype=torch.bfloat16`` (env ``SC_DP_GRAD=bf16``) sends the gradients as bf16 (88 MB instead of 177 MB at
    configs[2]) and writes the averaged result back into the fp32 gradients; the default (None) exchanges them in
    their own dtype (fp32).
    Gradients travel through ONE persistent flat buffer per bucket (no ``torch.cat`` per step); on CUDA the pack and
    unpack are one multi-tensor kernel launch each (``sc_grads_pack_multi`` / ``sc_grads_unpack_multi``).
    """

    def __init__(self, module: nn.Module, process_group=None, bucket_mb: Optional[float] = None,
                 average: bool = True, overlap: Optional[bool] = None, grad_dtype: Optional[torch.dtype] = None):
        super().__init__()
        self.module = module
        if overlap is None:
            overlap = os.environ.get("SC_DP_OVERLAP", "0") == "1"
        self.overlap = overlap
        if grad_dtype is None and os.environ.get("SC_DP_GRAD", "").lower() == "bf16":
            grad_dtype = torch.bfloat16
        if grad_dtype not in (None, torch.float32, torch.bfloat16):
            raise ValueError("grad_dtype must be None (the gradients' own dtype), torch.float32 or torch.bfloat16")
        self.grad_dtype = grad_dtype
        self.pg = process_group
        self.average = average
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.buckets: List[_Bucket] = []
        self._p2b = {}
        self._armed = False
        self.require_sync = True
        self.n_allreduce = 0                     # launches in the last backward (for tests/bench)
        if bucket_mb is None:                    # nothing to overlap with -> one collective for the whole model
            bucket_mb = 32.0 if overlap else float("inf")
        cap = bucket_mb * (1 << 20)
        cur, size = [], 0
        for p in reversed([p for p in module.parameters() if p.requires_grad]):
            cur.append(p)
            size += p.numel() * 4
            if size >= cap:
                self.buckets.append(_Bucket(cur))
                cur, size = [], 0
        if cur:
            self.buckets.append(_Bucket(cur))
        for bi, b in enumerate(self.buckets):
            b.offsets, b.total = _aligned_offsets(b.params)
            for p in b.params:
                self._p2b[p] = bi
                p.register_post_accumulate_grad_hook(self._hook)

    # -- forward is a pure pass-through: sharding happens in the data the rank feeds ------
    def forward(self, *a, **kw):
        return self.module(*a, **kw)

    def no_sync(self):
        """Context manager: accumulate locally (gradient accumulation steps, train.py:549)."""
        outer = self

        class _Ctx:
            def __enter__(self):
                outer.require_sync = False

            def __exit__(self, *exc):
                outer.require_sync = True
        return _Ctx()

    def _hook(self, p: nn.Parameter):
        if self.world == 1 or not self.require_sync:
            return
        if not self._armed:
            self._armed = True
            self.n_allreduce = 0
            for b in self.buckets:
                b.pending = len(b.params)
            torch.autograd.Variable._execution_engine.queue_callback(self._finish)
        b = self.buckets[self._p2b[p]]
        b.pending -= 1
        if b.pending == 0 and self.overlap:
            self._launch(b)

    @staticmethod
    def _native(g: torch.Tensor) -> bool:
        return g.is_cuda and g.dtype == torch.float32 and g.is_contiguous()

    def _launch(self, b: _Bucket):
        live = [(p, off) for p, off in zip(b.params, b.offsets) if p.grad is not None]
        if not live:
            return
        dev = live[0][0].grad.device
        fdt = self.grad_dtype or live[0][0].grad.dtype
        if b.flat is None or b.flat.device != dev or b.flat.dtype != fdt:
            b.flat = torch.zeros(b.total, dtype=fdt, device=dev)
        esz = b.flat.element_size()
        grads = [p.grad for p, _ in live]
        if all(self._native(g) for g in grads):
            from . import _lib
            from .optim import _counts, _table
            slices = (ctypes.c_void_p * len(live))(*[b.flat.data_ptr() + off * esz for _, off in live])
            with _lib.device_ctx(grads[0]):
                _lib.call("sc_grads_pack_multi", _table(grads), slices, _counts(grads), len(grads), _lib.dt(b.flat), _lib.stream())
        else:                                    # CPU tensors (gloo tests) / unusual layouts
            for (p, off), g in zip(live, grads):
                b.flat[off:off + g.numel()].copy_(g.reshape(-1))
        # NCCL averages inside the collective; gloo (CPU tests) only sums
        b.avg_in_collective = self.average and dist.get_backend(self.pg) == "nccl"
        op = dist.ReduceOp.AVG if b.avg_in_collective else dist.ReduceOp.SUM
        b.ev0 = None
        if dev.type == "cuda":
            from . import _lib as _l
            if _l.profile is not None:           # bench.py's timed region: bracket the collective with events on the compute stream
                b.ev0 = torch.cuda.Event(enable_timing=True)
                b.ev0.record()
        b.work = dist.all_reduce(b.flat, op=op, group=self.pg, async_op=True)
        b.live = live
        self.n_allreduce += 1

    def _finish(self):
        # params whose grad never arrived this backward (unused: the dead W_r / layernorm_r of
        # lucyrnn.py:50/56) leave their bucket incomplete on every rank alike: flush it now.
        for b in self.buckets:
            if b.work is None and b.pending != len(b.params) and (b.pending > 0 or not self.overlap):
                self._launch(b)
        for b in self.buckets:
            if b.work is None:
                continue
            b.work.wait()
            if getattr(b, "ev0", None) is not None:
                from . import _lib as _l
                ev1 = torch.cuda.Event(enable_timing=True)
                ev1.record()                     # after the compute stream has been made to wait for the collective:
                if _l.profile is not None:       # the collective itself + the wait for the slowest rank to enter it
                    _l.profile.append(("nccl_all_reduce", 0.0, b.ev0, ev1, ()))
                b.ev0 = None
            scale = 1.0 / self.world if (self.average and not b.avg_in_collective) else 1.0
            grads = [p.grad for p, _ in b.live]
            esz = b.flat.element_size()
            if all(self._native(g) for g in grads):
                from . import _lib
                from .optim import _counts, _table
                slices = (ctypes.c_void_p * len(b.live))(*[b.flat.data_ptr() + off * esz for _, off in b.live])
                with _lib.device_ctx(grads[0]):
                    _lib.call("sc_grads_unpack_multi", _table(grads), slices, _counts(grads), len(grads), _lib.dt(b.flat),
                              float(scale), _lib.stream())
            else:
                for (p, off), g in zip(b.live, grads):
                    g.copy_((b.flat[off:off + g.numel()].to(g.dtype) * scale).view_as(g))
            b.work, b.live = None, []
        self._armed = False

    def state_dict(self, *a, **kw):             # checkpoints interchange with the bare module
        return self.module.state_dict(*a, **kw)

    def load_state_dict(self, *a, **kw):
        return self.module.load_state_dict(*a, **kw)
