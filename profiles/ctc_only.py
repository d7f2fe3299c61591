#!/usr/bin/env python
"""CTC head alone at the cfg2 shape (B=64, T=3000, V=1024, U in [75,150]) — for ncu."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402
import statecatcher_b200 as sb  # noqa: E402

W = bench.WORKLOADS["cfg2"]
_, tok, inl, tgl = bench.synth_batch(W, 1234)
g = torch.Generator(device="cuda").manual_seed(0)
x = (torch.randn(W["B"], W["T"], W["V"], generator=g, device="cuda") * 2).bfloat16().requires_grad_(True)
tok = tok.cuda()
inl, tgl = torch.tensor(inl).cuda(), torch.tensor(tgl).cuda()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    x.grad = None
    loss = sb.ctc_loss_from_logits(x, tok, inl, tgl, zero_infinity=True)
    loss.backward()
torch.cuda.synchronize()
print("ok", loss.item())
