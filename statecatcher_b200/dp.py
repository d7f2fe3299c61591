"""Data parallelism by independent audio streams.

The reference has no distributed code at all; the only parallel axis the path offers is the
batch of streams: nothing inside LucyRNN (row-wise in b, lucyrnn.py:44-70) or CTC
(per-utterance lattice) mixes streams.  So each rank owns a contiguous block of streams and
the carried ``(h, s)`` of exactly those streams for the whole recording — state never crosses
ranks — and the single collective of a training step is a gradient all-reduce.

``StreamDataParallel`` buckets parameters in reverse registration order (= the order the
backward pass finishes them: output_proj first, layer 0 last), launches one asynchronous
NCCL all-reduce per bucket from a post-accumulate-grad hook the moment the bucket's last
gradient lands (so communication over NVLink overlaps the rest of the backward), and joins
all buckets in an end-of-backward callback.  Averaging by world size reproduces
``reduction='mean'`` over the GLOBAL batch when shards are equal (mean of rank means); for
unequal shards multiply the local loss by ``shard_loss_scale`` first.
"""
from __future__ import annotations

from typing import List, Optional

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
        self.flat: Optional[torch.Tensor] = None
        self.avg_in_collective = False


class StreamDataParallel(nn.Module):
    def __init__(self, module: nn.Module, process_group=None, bucket_mb: float = 32.0,
                 average: bool = True, overlap: Optional[bool] = None):
        super().__init__()
        self.module = module
        if overlap is None:
            overlap = os.environ.get("SC_DP_OVERLAP", "1") != "0"
        # overlap=False: every bucket is launched from the end-of-backward callback instead of
        # from its hook, so no NCCL kernel shares the SMs with the persistent GEMMs
        self.overlap = overlap
        self.pg = process_group
        self.average = average
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.buckets: List[_Bucket] = []
        self._p2b = {}
        self._armed = False
        self.require_sync = True
        self.n_allreduce = 0                     # launches in the last backward (for tests/bench)
        cap = int(bucket_mb * (1 << 20))
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

    def _launch(self, b: _Bucket):
        grads = [p.grad for p in b.params if p.grad is not None]
        if not grads:
            return
        b.flat = torch.cat([g.reshape(-1) for g in grads])
        # NCCL averages inside the collective; gloo (CPU tests) only sums
        b.avg_in_collective = self.average and dist.get_backend(self.pg) == "nccl"
        op = dist.ReduceOp.AVG if b.avg_in_collective else dist.ReduceOp.SUM
        b.work = dist.all_reduce(b.flat, op=op, group=self.pg, async_op=True)
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
            flat = b.flat
            if self.average and not b.avg_in_collective:
                flat = flat / self.world
            grads = [p.grad for p in b.params if p.grad is not None]
            views, off = [], 0
            for g in grads:
                views.append(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
            torch._foreach_copy_(grads, views)               # one fused copy-back per bucket
            b.work, b.flat = None, None
        self._armed = False

    def state_dict(self, *a, **kw):             # checkpoints interchange with the bare module
        return self.module.state_dict(*a, **kw)

    def load_state_dict(self, *a, **kw):
        return self.module.load_state_dict(*a, **kw)
