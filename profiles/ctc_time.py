#!/usr/bin/env python
"""Per-kernel-group timing of the CTC head at the cfg2 shape (CUDA events around the two C-ABI calls)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402
import statecatcher_b200 as sb  # noqa: E402
from statecatcher_b200 import _lib  # noqa: E402

W = dict(bench.WORKLOADS["cfg2"])
if os.environ.get("CTC_UMAX"):                 # lattice width study: transcripts of umax/2 .. umax labels
    W["umax"] = int(os.environ["CTC_UMAX"]); W["umin"] = max(1, W["umax"] - 2)
_, tok, inl, tgl = bench.synth_batch(W, 1234)
g = torch.Generator(device="cuda").manual_seed(0)
x = (torch.randn(W["B"], W["T"], W["V"], generator=g, device="cuda") * 2).bfloat16().requires_grad_(True)
tok = tok.cuda()
inl, tgl = torch.tensor(inl).cuda(), torch.tensor(tgl).cuda()
for it in range(8):
    if it == 3:
        _lib.profile = []
    x.grad = None
    loss = sb.ctc_loss_from_logits(x, tok, inl, tgl, zero_infinity=True)
    loss.backward()
torch.cuda.synchronize()
acc = {}
for name, _w, e0, e1, _s in _lib.profile:
    acc.setdefault(name, []).append(e0.elapsed_time(e1))
for k, v in acc.items():
    print(f"{k}: {sum(v) / len(v):.4f} ms (n={len(v)})")
print("loss", loss.item())
