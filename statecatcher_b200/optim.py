"""Fused gradient clipping + Adam/AdamW/Lion for the parameters of the path.

The reference clips with ``torch.nn.utils.clip_grad_norm_(model.parameters(), 50)``
(train.py:553), optionally logs the norm with one ``.item()`` sync per parameter
(train.py:555-560) and steps ``optim.Adam`` / ``optim.AdamW`` (train.py:112-137, 563-566).
``FusedAdam`` does all three with two kernels per tensor and no host synchronisation: the
global norm stays on the device (``.grad_norm`` is a 0-dim device tensor) and the update kernel
applies the clip coefficient on the fly.  ``clip_grad_norm_`` is the standalone clip with the
signature and return value of torch's.  ``Lion`` is the third choice of train.py's
``--optimizer`` (train.py:125-131 builds ``lion_pytorch.Lion(params, lr=, weight_decay=)``; that
package is absent and unpinned upstream, so the published update rule is what is implemented,
with the package's constructor signature and defaults).
"""
from __future__ import annotations

import ctypes

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


def _table(tensors):
    """HOST array of device pointers for the multi-tensor calls."""
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _counts(tensors):
    return (ctypes.c_int64 * len(tensors))(*[t.numel() for t in tensors])


def _global_sumsq(pgs, device, multi_tensor=False):
    acc = torch.zeros((), dtype=torch.float64, device=device)
    if multi_tensor:
        gs = [g for _, g in pgs]
        call("sc_sumsq_accum_multi", _table(gs), _counts(gs), len(gs), ptr(acc), stream())
        return acc
    for _, g in pgs:
        call("sc_sumsq_accum", ptr(g), g.numel(), ptr(acc), stream())
    return acc


def _check_param(p):
    if p.dtype != torch.float32 or not p.is_contiguous():
        raise TypeError("fused optimizer expects contiguous fp32 parameters")


def clip_grad_norm_(parameters, max_norm: float) -> torch.Tensor:
    """In-place clip of the global L2 norm; returns the pre-clip norm as a 0-dim DEVICE tensor."""
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    pgs = _grads(list(parameters))
    if not pgs:
        return torch.zeros(())
    with _lib.device_ctx(pgs[0][1]):
        acc = _global_sumsq(pgs, pgs[0][1].device)
        for _, g in pgs:
            call("sc_scale_grads", ptr(g), g.numel(), ptr(acc), float(max_norm), stream())
        return acc.sqrt().float()


class FusedAdam(torch.optim.Optimizer):
    """Adam (``decoupled=False``, weight decay as L2 like optim.Adam) or AdamW
    (``decoupled=True``) with optional fused global-norm clipping (``max_grad_norm``)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=True,
                 max_grad_norm=None, multi_tensor=True):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=decoupled))
        self.max_grad_norm = max_grad_norm
        self.grad_norm = None              # 0-dim device tensor after step() when clipping is on
        # one launch per 32 tensors (sc_*_multi) instead of one per tensor; same arithmetic per element.
        # Default since r02: AdamW + clip over the cfg2 parameter set (44.2 M in 26 tensors) 0.57 -> 0.30 ms on a B200
        # (profiles/r02_optim_time.txt); multi_tensor=False keeps one launch per tensor.
        self.multi_tensor = multi_tensor

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():                      # a standard closure calls loss.backward()
                loss = closure()
        allp = [p for grp in self.param_groups for p in grp["params"]]
        pgs = _grads(allp)
        if not pgs:
            return loss
        with _lib.device_ctx(pgs[0][1]):          # the C-ABI launches on the current device
            return self._step(pgs, loss)

    def _step(self, pgs, loss):
        acc = None
        if self.max_grad_norm is not None:
            acc = _global_sumsq(pgs, pgs[0][1].device, self.multi_tensor)
            self.grad_norm = acc.sqrt().float()
        for grp in self.param_groups:
            b1, b2 = grp["betas"]
            by_step = {}                                   # multi_tensor: tensors that share a step count
            for p in grp["params"]:
                if p.grad is None:
                    continue
                _check_param(p)
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p)
                    st["exp_avg_sq"] = torch.zeros_like(p)
                st["step"] += 1
                if self.multi_tensor:
                    by_step.setdefault(int(st["step"]), []).append(p)
                    continue
                call("sc_adam_step", ptr(p), ptr(p.grad), ptr(st["exp_avg"]), ptr(st["exp_avg_sq"]), p.numel(),
                     float(grp["lr"]), float(b1), float(b2), float(grp["eps"]), float(grp["weight_decay"]),
                     int(st["step"]), ptr(acc), float(self.max_grad_norm or 0.0), int(bool(grp["decoupled"])), stream())
            for step, ps in by_step.items():
                call("sc_adam_step_multi", _table(ps), _table([p.grad for p in ps]),
                     _table([self.state[p]["exp_avg"] for p in ps]), _table([self.state[p]["exp_avg_sq"] for p in ps]),
                     _counts(ps), len(ps), float(grp["lr"]), float(b1), float(b2), float(grp["eps"]),
                     float(grp["weight_decay"]), step, ptr(acc), float(self.max_grad_norm or 0.0),
                     int(bool(grp["decoupled"])), stream())
        return loss


class Lion(torch.optim.Optimizer):
    """``lion_pytorch.Lion(params, lr=1e-4, betas=(0.9, 0.99), weight_decay=0.0)`` (the call of
    train.py:125-131) as one kernel per tensor: decoupled decay, ``p -= lr * sign(b1*m + (1-b1)*g)``,
    ``m = b2*m + (1-b2)*g``; ``max_grad_norm`` fuses the global-norm clip like ``FusedAdam``."""

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.99), weight_decay=0.0, max_grad_norm=None, multi_tensor=True):
        if lr <= 0.0:
            raise ValueError("lr must be positive")
        if not all(0.0 <= b <= 1.0 for b in betas):
            raise ValueError("betas must lie in [0, 1]")
        super().__init__(params, dict(lr=lr, betas=betas, weight_decay=weight_decay))
        self.max_grad_norm = max_grad_norm
        self.grad_norm = None
        self.multi_tensor = multi_tensor   # see FusedAdam

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        allp = [p for grp in self.param_groups for p in grp["params"]]
        pgs = _grads(allp)
        if not pgs:
            return loss
        with _lib.device_ctx(pgs[0][1]):
            return self._step(pgs, loss)

    def _step(self, pgs, loss):
        acc = None
        if self.max_grad_norm is not None:
            acc = _global_sumsq(pgs, pgs[0][1].device, self.multi_tensor)
            self.grad_norm = acc.sqrt().float()
        for grp in self.param_groups:
            b1, b2 = grp["betas"]
            batch = []
            for p in grp["params"]:
                if p.grad is None:
                    continue
                _check_param(p)
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p)
                if self.multi_tensor:
                    batch.append(p)
                    continue
                call("sc_lion_step", ptr(p), ptr(p.grad), ptr(st["exp_avg"]), p.numel(), float(grp["lr"]),
                     float(b1), float(b2), float(grp["weight_decay"]), ptr(acc), float(self.max_grad_norm or 0.0),
                     stream())
            if batch:
                call("sc_lion_step_multi", _table(batch), _table([p.grad for p in batch]),
                     _table([self.state[p]["exp_avg"] for p in batch]), _counts(batch), len(batch), float(grp["lr"]),
                     float(b1), float(b2), float(grp["weight_decay"]), ptr(acc), float(self.max_grad_norm or 0.0),
                     stream())
        return loss
