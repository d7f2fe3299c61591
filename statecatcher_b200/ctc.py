"""CTC loss on the sm_100a kernels (K3) — drop-in for ``torch.nn.CTCLoss``.

The reference builds ``nn.CTCLoss(blank=blank_id, zero_infinity=True)`` (train.py:142) and
calls it as ``criterion(enc_out.log_softmax(-1).transpose(0,1), tokens, in_lens, tgt_lens)``
(model.py:70-71) with Python lists for the lengths.  ``CTCLoss`` here takes the same
positional arguments and returns a 0-dim tensor with grad.  The kernels fold log-softmax in
(it is idempotent), so the input may be raw logits or already-normalised log-probs, in
(T,B,V) layout with arbitrary T/B strides (the transposed view of a (B,T,V) tensor is read in
place).  The gradient returned for the input is ``scale*(softmax - occupancy)`` — identical
to what ATen returns for normalised input, and equal to d loss/d logits.
Semantics: 'mean' = mean_b(nll_b/max(U_b,1)); infeasible utterances
(+inf) give loss 0 and an all-zero gradient row when zero_infinity=True; grad is exactly 0
for frames t >= T_b.
"""
from __future__ import annotations

import os
from typing import Sequence, Union

import torch
import torch.nn as nn

from . import _lib
from ._lib import call, dt, ptr, stream

_RED = {"none": 0, "mean": 1, "sum": 2}
LenT = Union[Sequence[int], torch.Tensor]

# The overlapped head (sc_ctc_head), opt-in with SC_CTC_OVERLAP=1: when the loss is a scalar and its input wants a
# gradient, the forward forms loss AND gradient in one call, the two V-wide passes running on side streams under the
# latency-bound recursions; the backward then only multiplies by the upstream gradient (a no-op when that is 1, the
# `loss.backward()` of train.py:549).  Bit-identical to the three passes (tests/test_gpu_ctc_head.py), but measured
# on a B200 at the cfg2 shape it does not pay yet (profiles/r02_ctc_head_exp.txt: 0.79-0.89 ms against 0.81 ms alone,
# 1.12 against 1.04 ms inside the step): a frame's gradient row needs both directions, so the gradient pass can only
# start at the half-way point, and next to a recursion block (164 KB of shared memory) only one of its blocks fits on
# an SM.  Default: the three passes one after the other, gradient in the backward.
_OVERLAP = os.environ.get("SC_CTC_OVERLAP", "0") == "1"
_side_streams = {}


def _head_streams(device):
    """Handles of the three side streams of ``device`` (created once): a high-priority one for the recursion
    launches, two of default priority for the emission and the gradient pass.  (0, 0, 0) — which sc_ctc_head takes
    as "one pass after the other" — for a tensor that is not on a CUDA device (only the host-logic tests get here:
    ``require_cuda`` has already refused such a tensor)."""
    if device.type != "cuda":
        return 0, 0, 0
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _side_streams:
        _side_streams[key] = (torch.cuda.Stream(device=key, priority=-1), torch.cuda.Stream(device=key),
                              torch.cuda.Stream(device=key))
    return tuple(s.cuda_stream for s in _side_streams[key])


def _lplat_pitch(Umax: int, S: int) -> int:
    """Words per frame of the emission workspace (sc_ctc_lplat_pitch; S when the library is not there: host tests)."""
    try:
        return int(_lib.load().sc_ctc_lplat_pitch(Umax))
    except (ImportError, OSError):
        return S


def _lens(v: LenT, device, B: int, name: str):
    """-> (int64 device tensor [B], host max or None)."""
    if isinstance(v, torch.Tensor):
        t = v.to(device=device, dtype=torch.int64).contiguous()
        hmax = None
    else:
        vals = [int(a) for a in v]
        t = torch.tensor(vals, dtype=torch.int64, device=device)
        hmax = max(vals) if vals else 0
    if t.numel() != B:
        raise ValueError(f"{name} must have batch size {B} entries, got {t.numel()}")
    return t, hmax


class _CTCFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, targets, in_lens, tgt_lens, blank, red, Umax):
        # x: logical (T,B,V), V contiguous
        T, B, V = x.shape
        dev = x.device
        S = (2 * Umax + 1 + 7) & ~7          # lattice row width, padded to whole 32-byte sectors
        f32 = dict(dtype=torch.float32, device=dev)
        lse = torch.empty(B, max(T, 1), **f32)
        lplat = torch.empty(B, max(T, 1), _lplat_pitch(Umax, S), **f32)   # emission rows (>= S words per frame)
        cshift = torch.empty(B, max(T, 1), **f32)
        alpha = torch.empty(B, max(T, 1), S, **f32)
        beta = torch.empty(B, max(T, 1), S, **f32)
        nll = torch.empty(B, **f32)
        loss = torch.zeros((), **f32)
        # opaque per-call workspace of the lattice pass (format flags, per-direction likelihoods, range records)
        ws = torch.empty(_lib.load().sc_ctc_workspace_bytes(B, T, Umax) // 8 + 1, dtype=torch.float64, device=dev)
        ldt = targets.stride(0) if targets.numel() else max(Umax, 1)
        ctx.dx = None
        if _OVERLAP and red != 0 and T > 0 and ctx.needs_input_grad[0]:
            dx = _CTCFn._grad_like(x)
            sl, se, sg = _head_streams(dev)
            call("sc_ctc_head", ptr(x), x.stride(1), x.stride(0), dt(x), ptr(targets), ldt,
                 ptr(in_lens), ptr(tgt_lens), B, T, V, Umax, blank, ptr(lse), ptr(lplat), ptr(cshift),
                 ptr(alpha), ptr(beta), ptr(nll), ptr(loss), red, ptr(ws),
                 ptr(dx), dx.stride(1), dx.stride(0), dt(dx), 0, stream(), sl, se, sg)
            ctx.dx = dx                      # for a unit upstream gradient; the backward scales it
        else:
            call("sc_ctc_emissions", ptr(x), x.stride(1), x.stride(0), dt(x), ptr(targets), ldt,
                 ptr(in_lens), ptr(tgt_lens), B, T, V, Umax, blank, ptr(lse), ptr(lplat), ptr(cshift), stream())
            call("sc_ctc_lattice", ptr(lplat), ptr(cshift), ptr(targets), ldt, ptr(in_lens), ptr(tgt_lens), B, T, Umax, blank,
                 ptr(alpha), ptr(beta), ptr(nll), ptr(loss), red, ptr(ws), stream())
        ctx.save_for_backward(x, targets, in_lens, tgt_lens, lse, alpha, beta, nll, ws)
        ctx.cfg = (blank, red, Umax, ldt)
        if red != 0:
            ctx.mark_non_differentiable(nll)     # per-utterance nll is a by-product here
        return loss, nll

    @staticmethod
    def _grad_like(x):
        dense = x.is_contiguous() or x.transpose(0, 1).is_contiguous()
        return torch.empty_strided(x.shape, x.stride(), dtype=x.dtype, device=x.device) if dense \
            else torch.empty_like(x, memory_format=torch.contiguous_format)

    @staticmethod
    def backward(ctx, gout, _gnll):
        x, targets, in_lens, tgt_lens, lse, alpha, beta, nll, ws = ctx.saved_tensors
        blank, red, Umax, ldt = ctx.cfg
        T, B, V = x.shape
        g = (_gnll if red == 0 else gout).to(torch.float32).contiguous()
        if ctx.dx is not None:               # formed during the forward; a second backward over a retained graph recomputes
            dx, ctx.dx = ctx.dx, None
            call("sc_ctc_scale_grad", ptr(dx), dt(dx), dx.numel(), ptr(g), stream())
            return dx, None, None, None, None, None, None
        dx = _CTCFn._grad_like(x)
        call("sc_ctc_bwd", ptr(x), x.stride(1), x.stride(0), dt(x), ptr(targets), ldt,
             ptr(in_lens), ptr(tgt_lens), B, T, V, Umax, blank, ptr(lse), ptr(alpha), ptr(beta),
             ptr(nll), ptr(g), red, ptr(dx), dx.stride(1), dx.stride(0), dt(dx), ptr(ws), stream())
        return dx, None, None, None, None, None, None


@_lib.on_tensor_device
def ctc_loss(log_probs: torch.Tensor, targets: torch.Tensor, input_lengths: LenT,
             target_lengths: LenT, blank: int = 0, reduction: str = "mean",
             zero_infinity: bool = False) -> torch.Tensor:
    """Signature of ``torch.nn.functional.ctc_loss``; input (T,B,V) or (T,V)."""
    if reduction not in _RED:
        raise ValueError(f"{reduction} is not a valid value for reduction")
    _lib.require_cuda(log_probs, "ctc_loss input")
    x = log_probs
    unbatched = x.dim() == 2
    if unbatched:
        x = x.unsqueeze(1)
        if targets.dim() == 1 and not isinstance(target_lengths, torch.Tensor) and len(target_lengths) == 1:
            targets = targets.unsqueeze(0)
    if x.dim() != 3:
        raise ValueError("ctc_loss expects (T,B,V) or (T,V) input")
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    if x.size(2) > 1 and x.stride(2) != 1:
        x = x.contiguous()
    T, B, V = x.shape
    in_t, _ = _lens(input_lengths, x.device, B, "input_lengths")
    tg_t, umax_host = _lens(target_lengths, x.device, B, "target_lengths")
    tg = targets.to(device=x.device, dtype=torch.int64)
    if tg.dim() == 1:                                  # concatenated targets -> padded
        lens_host = tg_t.tolist()
        Um = max(lens_host) if lens_host else 0
        padded = torch.full((B, max(Um, 1)), blank, dtype=torch.int64, device=x.device)
        off = 0
        for b, u in enumerate(lens_host):
            padded[b, :u] = tg[off:off + u]
            off += u
        tg, umax_host = padded, Um
    tg = tg.contiguous()
    Umax = tg.size(1) if tg.dim() == 2 else 0
    if umax_host is not None:
        if umax_host > Umax:
            raise ValueError("target_lengths exceed the targets tensor width")
        Umax = umax_host                               # smaller lattice workspace
    loss, nll = _CTCFn.apply(x, tg, in_t, tg_t, int(blank), _RED[reduction], int(Umax))
    if reduction == "none":
        out = nll
        if zero_infinity:
            out = torch.where(torch.isinf(out), torch.zeros_like(out), out)
        return out
    if not zero_infinity:
        # kernels implement the zero_infinity reduction; restore +inf for the strict variant
        bad = torch.isinf(nll).any()
        loss = torch.where(bad, torch.full_like(loss, float("inf")), loss)
    return loss


def ctc_loss_from_logits(logits_btv: torch.Tensor, targets, input_lengths, target_lengths,
                         blank: int = 0, reduction: str = "mean", zero_infinity: bool = True):
    """Fused head: (B,T,V) encoder output in, loss out (model.py:70-71 in one call)."""
    return ctc_loss(logits_btv.transpose(0, 1), targets, input_lengths, target_lengths,
                    blank, reduction, zero_infinity)


class CTCLoss(nn.Module):
    """``torch.nn.CTCLoss``-compatible module running the K3 kernels."""

    def __init__(self, blank: int = 0, reduction: str = "mean", zero_infinity: bool = False):
        super().__init__()
        self.blank, self.reduction, self.zero_infinity = blank, reduction, zero_infinity

    def forward(self, log_probs, targets, input_lengths, target_lengths):
        return ctc_loss(log_probs, targets, input_lengths, target_lengths, self.blank,
                        self.reduction, self.zero_infinity)
