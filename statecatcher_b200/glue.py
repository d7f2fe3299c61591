"""Host glue around the hot path, mirroring the reference's model.py.

* ``detach_states`` / ``assert_all_detached``  — model.py:11-35 (bit-exact state handoff:
  tensors are detached, never copied or re-typed).
* ``compute_loss``  — model.py:37-110: detach carried state, run the model, apply the loss
  head, return the reference's 4-tuple ``(loss, state, enc_out, state)``.
* ``LucyASRModel``  — the LucyRNN branch of ``ASRModel`` (model.py:282-398): optional input
  projection, zero-masking of padded frames (model.py:376-377), positional encoder call
  (model.py:384/388).  The LSTM/xLSTM branches are other encoders and out of scope.
"""
from __future__ import annotations

from typing import Any, Optional

import torch
import torch.nn as nn

from . import ops
from .ctc import CTCLoss, ctc_loss_from_logits
from .lucyrnn import LucyRNN
from .lucyrnn_conf import LucyRNNConfig


def detach_states(states):
    if states is None:
        return None
    if isinstance(states, torch.Tensor):
        return states.detach()
    if isinstance(states, dict):
        return {k: detach_states(v) for k, v in states.items()}
    if isinstance(states, tuple):
        return tuple(detach_states(v) for v in states)
    if isinstance(states, list):
        return [detach_states(v) for v in states]
    return states


def assert_all_detached(x):
    if isinstance(x, torch.Tensor):
        assert not x.requires_grad, "Tensor still requires grad"
    elif isinstance(x, (list, tuple)):
        for v in x:
            assert_all_detached(v)
    elif isinstance(x, dict):
        for v in x.values():
            assert_all_detached(v)


class LucyASRModel(nn.Module):
    def __init__(self, cfg: LucyRNNConfig, frontend: Optional[nn.Module] = None, feat_dim: int = 80,
                 proj_dim: int = -1, compute_dtype=None):
        super().__init__()
        self.frontend = frontend
        self.cfg = cfg                      # train.py:484 reads model.cfg.stack_order
        self.encoder = LucyRNN(cfg, compute_dtype=compute_dtype)
        self.enc_out_dim = cfg.vocab_size
        if proj_dim > 0:
            self.proj = nn.Linear(feat_dim, proj_dim)

    def forward(self, feats, mask, states=None):
        if hasattr(self, "proj"):
            feats = self.proj(feats)
        if mask is not None:
            if feats.requires_grad or feats.dtype not in (torch.float32, torch.bfloat16):
                feats = feats * mask.unsqueeze(-1).float()       # keeps autograd through an input projection
            else:
                feats = ops.mask_rows(feats, mask)               # model.py:377, our kernel
        if states is not None:
            return self.encoder(feats, states)
        return self.encoder(feats)


def compute_loss(mode: str, criterion, model: nn.Module, feats, masks, tokens, in_lens, tgt_lens,
                 blank_id: int, use_rnnt_joiner: Optional[nn.Module] = None,
                 input_state: Optional[Any] = None, args=None, compact=False):
    if input_state:                                          # truthiness gate, model.py:60
        input_state = detach_states(input_state)
        if args and getattr(args, "debug", False):
            assert_all_detached(input_state)
    enc_out, output_state = model(feats, masks, input_state)
    if mode == "ctc":
        if isinstance(criterion, CTCLoss):
            # fused head: log-softmax folded into the CTC kernels, (B,T,V) read in place
            loss = ctc_loss_from_logits(enc_out, tokens, in_lens, tgt_lens, criterion.blank,
                                        criterion.reduction, criterion.zero_infinity)
        else:
            loss = criterion(enc_out.log_softmax(-1).transpose(0, 1), tokens, in_lens, tgt_lens)
    elif mode == "rnnt":
        assert use_rnnt_joiner is not None, "Joiner module required for RNN-T mode"
        prefix = torch.cat([torch.full((tokens.size(0), 1), blank_id, dtype=tokens.dtype,
                                       device=tokens.device), tokens], dim=1)
        if args is not None and getattr(args, "compact_rnnt", False):
            logits = use_rnnt_joiner(enc_out, prefix, in_lens, tgt_lens)
        else:
            logits = use_rnnt_joiner(enc_out, prefix)
        log_probs = logits.float().log_softmax(dim=-1)
        loss = criterion(log_probs=log_probs, labels=tokens, frames_lengths=in_lens,
                         labels_lengths=tgt_lens, blank_id=blank_id, compact=compact, gather=True)
    else:
        raise ValueError(f"Unknown mode: {mode}")
    return loss, output_state, enc_out, output_state


class SegmentPrefetcher:
    """Host -> device staging of segment batches, one batch ahead of the compute stream.

    The reference moves each batch with a blocking ``.to(device)`` right before the forward
    (train.py:505-512); at B200 speed that copy (61 MB of fp32 features for 64 x 3000 frames)
    is ~3 % of a step.  This feeder issues the copy of batch i+1 on its own CUDA stream while
    batch i computes, into one of two device buffers, and hands the consumer tensors that are
    already ordered after the copy on the current stream.  Host tensors should be pinned.

        feeder = SegmentPrefetcher(iter_of_tuples_of_host_tensors, device)
        for feats, tokens, in_lens, tgt_lens in feeder: ...   # non-tensors pass through

    A buffer is reused two batches later; the copy stream waits for the compute stream's
    position at the time the *previous* batch was handed out, so a buffer is never overwritten
    while kernels still read it.
    """

    def __init__(self, batches, device):
        self.it = iter(batches)
        self.dev = torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.dev)
        self.bufs = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.released = [None, None]          # compute-stream events guarding buffer reuse
        self.slot = 0
        self.pending = None
        # the staging buffers come from the caching allocator of the CURRENT stream: a recycled block may
        # still be read by compute work already in flight, so the copy stream starts behind it
        self.copy_stream.wait_stream(torch.cuda.current_stream(self.dev))
        self._stage()

    def _stage(self):
        try:
            host = next(self.it)
        except StopIteration:
            self.pending = None
            return
        k = self.slot
        isT = [torch.is_tensor(h) for h in host]
        bufs = self.bufs[k]
        if bufs is None or len(bufs) != len(host) or any(
                t and (not torch.is_tensor(b) or b.shape != h.shape or b.dtype != h.dtype)
                for b, h, t in zip(bufs, host, isT)):
            bufs = [torch.empty(h.shape, dtype=h.dtype, device=self.dev) if t else None
                    for h, t in zip(host, isT)]
        with torch.cuda.stream(self.copy_stream):
            if self.released[k] is not None:
                self.copy_stream.wait_event(self.released[k])
            for i, (h, t) in enumerate(zip(host, isT)):
                if t:
                    bufs[i].copy_(h, non_blocking=True)
                else:
                    bufs[i] = h                       # lengths lists etc. pass through
            self.ready[k].record(self.copy_stream)
        self.bufs[k] = bufs
        self.pending = k
        self.slot ^= 1

    def __iter__(self):
        return self

    def __next__(self):
        if self.pending is None:
            raise StopIteration
        k = self.pending
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(self.ready[k])
        out = tuple(self.bufs[k])
        # the other buffer was handed out one batch ago: everything enqueued on the compute
        # stream up to now has finished with it once this event fires
        ev = torch.cuda.Event()
        ev.record(cur)
        self.released[k ^ 1] = ev
        self._stage()
        return out


class GraphedStreamingEncoder:
    """Streaming forward (``is_training=False`` step path) of a LucyRNN replayed from a CUDA graph.

    One segment of one live stream is ~60 small launches whose GPU work (0.7 ms at 6 x 1024,
    3000 frames) is shorter than the time Python needs to enqueue them; capturing the segment
    once and replaying it removes the launch path from the latency.  Input, logits and the
    carried ``(h, s)`` state live in static buffers; ``step(x)`` copies the segment in, replays,
    and returns the logits buffer (valid until the next ``step``) — the state of segment k seeds
    segment k+1 inside the graph, exactly as ``model(x, state)`` would chain it.

        runner = GraphedStreamingEncoder(model, batch=1, frames=3000, feat_dim=80)
        for seg in segments: logits = runner.step(seg)
    """

    def __init__(self, model: LucyRNN, batch: int, frames: int, feat_dim: int, device=None,
                 in_dtype: torch.dtype = torch.float32, warmup: int = 3):
        cfg = model.config
        if cfg.is_training:
            raise ValueError("GraphedStreamingEncoder needs a model built with is_training=False (step path)")
        if not cfg.return_last_states:
            raise ValueError("GraphedStreamingEncoder needs return_last_states=True")
        self.model = model
        dev = torch.device(device) if device is not None else next(model.parameters()).device
        H, L = cfg.hidden_dim, cfg.num_layers
        self.x = torch.zeros(batch, frames, feat_dim, device=dev, dtype=in_dtype)
        self.h = [torch.zeros(batch, H, device=dev) for _ in range(L)]
        self.s = [torch.zeros(batch, H, device=dev) for _ in range(L)]
        self.logits = None
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(warmup, 1)):                    # lazy one-time setup (func attributes, driver entry points)
                self._segment()
            self.reset()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.logits = self._segment()

    def _segment(self):
        logits, (h2, s2) = self.model(self.x, (list(self.h), list(self.s)))
        for dst, src in zip(self.h + self.s, list(h2) + list(s2)):
            dst.copy_(src)
        return logits

    def reset(self):
        """Start a new stream: zero the carried state."""
        for t in self.h + self.s:
            t.zero_()

    def step(self, x: torch.Tensor) -> torch.Tensor:
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.logits

    @property
    def state(self):
        return self.h, self.s


class GraphedTrainStep:
    """One training step of a LucyRNN — carried state in, forward, fused CTC loss, backward, carried state out —
    replayed from ONE CUDA graph.

    For small models (BASELINE configs[0]: 2 x 256, 8 streams x 1000 frames) a step is a chain of ~300 launches
    of a few microseconds each and the GPU waits for Python; captured once, the chain replays without the launch
    path.  Features, labels, lengths, the carried ``(h, s)`` and the loss live in static buffers; the parameter
    gradients are the graph's own tensors, OVERWRITTEN by every replay (no need to clear them; ``step``
    re-attaches them to ``p.grad`` when ``zero_grad(set_to_none=True)`` has dropped them).
    The state of segment k seeds segment k+1 inside the graph, detached at the boundary exactly as
    ``compute_loss`` does it (model.py:60-63); ``reset()`` starts new streams.

        runner = GraphedTrainStep(model, batch=8, frames=1000, feat_dim=80, max_labels=50)
        for feats, tokens, in_lens, tgt_lens in segments:
            loss = runner.step(feats, tokens, in_lens, tgt_lens)      # gradients are in p.grad
            optimizer.step()

    Single GPU (the hooks of ``StreamDataParallel`` cannot run inside a capture).  Written after round 1's GPU
    budget was spent: compiled into the package and covered by tests/test_gpu_zzz_graph_train.py, not yet timed.
    """

    def __init__(self, model: LucyRNN, batch: int, frames: int, feat_dim: int, max_labels: int, blank: int = 0,
                 reduction: str = "mean", zero_infinity: bool = True, device=None,
                 in_dtype: torch.dtype = torch.float32, warmup: int = 3):
        cfg = model.config
        if not cfg.is_training:
            raise ValueError("GraphedTrainStep needs a model built with is_training=True (segment path)")
        if not cfg.return_last_states:
            raise ValueError("GraphedTrainStep needs return_last_states=True")
        self.model = model
        self.head = (int(blank), reduction, bool(zero_infinity))
        dev = torch.device(device) if device is not None else next(model.parameters()).device
        H, L = cfg.hidden_dim, cfg.num_layers
        self.x = torch.zeros(batch, frames, feat_dim, device=dev, dtype=in_dtype)
        self.tokens = torch.zeros(batch, max(max_labels, 1), device=dev, dtype=torch.int64)
        self.in_lens = torch.full((batch,), frames // max(cfg.stack_order, 1), device=dev, dtype=torch.int64)
        self.tgt_lens = torch.zeros(batch, device=dev, dtype=torch.int64)
        self.h = [torch.zeros(batch, H, device=dev) for _ in range(L)]
        self.s = [torch.zeros(batch, H, device=dev) for _ in range(L)]
        self.loss = None
        params = [p for p in model.parameters() if p.requires_grad]
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):                    # lazy one-time setup outside the capture
                for p in params:
                    p.grad = None
                self._step()
            for p in params:
                p.grad = None                                  # the capture allocates the static gradients
            self.reset()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._step()
        # the captured gradient tensors are the only link between a replay and the optimizer
        self._params = params
        self._grads = [p.grad for p in params]

    def _step(self):
        blank, reduction, zero_infinity = self.head
        logits, (h2, s2) = self.model(self.x, (list(self.h), list(self.s)))
        loss = ctc_loss_from_logits(logits, self.tokens, self.in_lens, self.tgt_lens, blank, reduction, zero_infinity)
        loss.backward()
        with torch.no_grad():
            for dst, src in zip(self.h + self.s, list(h2) + list(s2)):
                if src is not dst:                             # training path hands the caller's s back untouched
                    dst.copy_(src.detach())
        return loss.detach()

    def reset(self):
        """Start new streams: zero the carried state."""
        for t in self.h + self.s:
            t.zero_()

    def step(self, feats: torch.Tensor, tokens: torch.Tensor, in_lens, tgt_lens) -> torch.Tensor:
        if tokens.dim() != 2 or tokens.size(0) != self.tokens.size(0) or tokens.size(1) > self.tokens.size(1):
            raise ValueError("tokens must be (batch, U) with U <= max_labels")
        self.x.copy_(feats, non_blocking=True)
        self.tokens.zero_()
        self.tokens[:, :tokens.size(1)].copy_(tokens, non_blocking=True)
        for dst, src, name in ((self.in_lens, in_lens, "in_lens"), (self.tgt_lens, tgt_lens, "tgt_lens")):
            t = src if torch.is_tensor(src) else torch.tensor([int(v) for v in src], dtype=torch.int64)
            if t.numel() != dst.numel():
                raise ValueError(f"{name} must have {dst.numel()} entries")
            dst.copy_(t.to(torch.int64), non_blocking=True)
        # optimizer.zero_grad() (set_to_none=True is torch's default) or a user assignment may have
        # detached p.grad from the graph's own gradient tensor: replay would still write the captured
        # buffer while the optimizer saw nothing and silently stopped updating.  Re-attach.
        for p, g in zip(self._params, self._grads):
            if p.grad is not g:
                p.grad = g
        self.graph.replay()
        return self.loss

    @property
    def state(self):
        return self.h, self.s
