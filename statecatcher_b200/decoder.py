"""Greedy CTC decoding on the GPU — drop-in for the reference's decoder.py.

``ctc_greedy_decoder(log_probs, input_lengths, blank=0) -> List[List[int]]`` keeps the
signature of decoder.py:3-30 (called from train.py:232 every 100 steps); the argmax and the
blank/repeat collapse run as two sm_100a kernels and the result comes back in ONE device->host
copy instead of one ``.item()`` sync per token.
"""
from __future__ import annotations

from typing import List

import torch

from . import _lib
from ._lib import call, dt, ptr, stream
from .ctc import _lens


@_lib.on_tensor_device
def ctc_greedy_decoder(log_probs: torch.Tensor, input_lengths, blank: int = 0) -> List[List[int]]:
    _lib.require_cuda(log_probs, "ctc_greedy_decoder input")
    x = log_probs
    if x.dim() != 3:
        raise ValueError("ctc_greedy_decoder expects (batch, time, vocab)")
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    if x.size(2) > 1 and x.stride(2) != 1:
        x = x.contiguous()
    B, T, V = x.shape
    lens, _ = _lens(input_lengths, x.device, B, "input_lengths")
    Tm = max(T, 1)
    pred = torch.empty(B, Tm, dtype=torch.int32, device=x.device)
    buf = torch.empty(B * Tm + B, dtype=torch.int64, device=x.device)          # tokens [B,T] then lengths [B]
    toks, olen = buf[:B * Tm].view(B, Tm), buf[B * Tm:]
    call("sc_ctc_greedy_decode", ptr(x), x.stride(0), x.stride(1), dt(x), ptr(lens), B, T, V, int(blank),
         ptr(pred), ptr(toks), ptr(olen), stream())
    host = buf.cpu()                                                            # the only device->host copy
    ht, hl = host[:B * Tm].view(B, Tm), host[B * Tm:]
    return [ht[b, :int(hl[b])].tolist() for b in range(B)]
