"""Fused gradient clipping + Adam/AdamW for the parameters of the path.

The reference clips with ``torch.nn.utils.clip_grad_norm_(model.parameters(), 50)``
(train.py:553), optionally logs the norm with one ``.item()`` sync per parameter
(train.py:555-560) and steps ``optim.Adam`` / ``optim.AdamW`` (train.py:112-137, 563-566).
``FusedAdam`` does all three with two kernels per tensor and no host synchronisation: the
global norm stays on the device (``.grad_norm`` is a 0-dim device tensor) and the update kernel
applies the clip coefficient on the fly.  ``clip_grad_norm_`` is the standalone clip with the
signature and return value of torch's.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, ptr, stream


def _grads(params):
    out = []
    for p in params:
        if p.grad is None:
            continue
        g = p.grad
        if g.dtype != torch.float32 or not g.is_contiguous():
            raise TypeError("fused optimizer expects contiguous fp32 gradients")
        _lib.require_cuda(g, "gradient")
        out.append((p, g))
    return out


def _global_sumsq(pgs, device):
    acc = torch.zeros((), dtype=torch.float64, device=device)
    for _, g in pgs:
        call("sc_sumsq_accum", ptr(g), g.numel(), ptr(acc), stream())
    return acc


def clip_grad_norm_(parameters, max_norm: float) -> torch.Tensor:
    """In-place clip of the global L2 norm; returns the pre-clip norm as a 0-dim DEVICE tensor."""
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    pgs = _grads(list(parameters))
    if not pgs:
        return torch.zeros(())
    acc = _global_sumsq(pgs, pgs[0][1].device)
    for _, g in pgs:
        call("sc_scale_grads", ptr(g), g.numel(), ptr(acc), float(max_norm), stream())
    return acc.sqrt().float()


class FusedAdam(torch.optim.Optimizer):
    """Adam (``decoupled=False``, weight decay as L2 like optim.Adam) or AdamW
    (``decoupled=True``) with optional fused global-norm clipping (``max_grad_norm``)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=True,
                 max_grad_norm=None):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=decoupled))
        self.max_grad_norm = max_grad_norm
        self.grad_norm = None              # 0-dim device tensor after step() when clipping is on

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        allp = [p for grp in self.param_groups for p in grp["params"]]
        pgs = _grads(allp)
        if not pgs:
            return loss
        acc = None
        if self.max_grad_norm is not None:
            acc = _global_sumsq(pgs, pgs[0][1].device)
            self.grad_norm = acc.sqrt().float()
        for grp in self.param_groups:
            b1, b2 = grp["betas"]
            for p in grp["params"]:
                if p.grad is None:
                    continue
                if p.dtype != torch.float32 or not p.is_contiguous():
                    raise TypeError("fused optimizer expects contiguous fp32 parameters")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                st["step"] += 1
                call("sc_adam_step", ptr(p), ptr(p.grad), ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]), p.numel(),
                     float(grp["lr"]), float(b1), float(b2), float(grp["eps"]), float(grp["weight_decay"]),
                     int(st["step"]), ptr(acc), float(self.max_grad_norm or 0.0), int(bool(grp["decoupled"])), stream())
        return loss
