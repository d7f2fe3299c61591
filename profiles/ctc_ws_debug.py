#!/usr/bin/env python
"""Debug aid: run the CTC forward at the cfg2 shape and print what the fp64 lattice's range checks recorded."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
from statecatcher_b200 import _lib
from statecatcher_b200._lib import call, dt, ptr, stream
W = bench.WORKLOADS["cfg2"]
_, tok, inl, tgl = bench.synth_batch(W, 1234)
B, T, V = W["B"], W["T"], W["V"]
g = torch.Generator(device="cuda").manual_seed(0)
x = (torch.randn(B, T, V, generator=g, device="cuda") * 2).bfloat16()
tok = tok.cuda(); U = tok.size(1)
il, tl = torch.tensor(inl).cuda(), torch.tensor(tgl).cuda()
S = (2 * U + 1 + 7) & ~7
f32 = dict(dtype=torch.float32, device="cuda")
lse, lplat, csh = torch.zeros(B, T, **f32), torch.zeros(B, T, _lib.load().sc_ctc_lplat_pitch(U), **f32), torch.zeros(B, T, **f32)
alpha, beta = torch.zeros(B, T, S, **f32), torch.zeros(B, T, S, **f32)
nll, loss = torch.zeros(B, **f32), torch.zeros((), **f32)
nb = _lib.load().sc_ctc_workspace_bytes(B, T, U)
ws = torch.zeros(nb // 8 + 1, dtype=torch.float64, device="cuda")
call("sc_ctc_emissions", ptr(x), x.stride(0), x.stride(1), dt(x), ptr(tok), tok.stride(0), ptr(il), ptr(tl), B, T, V, U, 0, ptr(lse), ptr(lplat), ptr(csh), stream())
call("sc_ctc_lattice", ptr(lplat), ptr(csh), ptr(tok), tok.stride(0), ptr(il), ptr(tl), B, T, U, 0, ptr(alpha), ptr(beta), ptr(nll), ptr(loss), 1, ptr(ws), stream())
torch.cuda.synchronize()
zl2 = ws[:2 * B].cpu().view(B, 2)
raw = ws.cpu().view(torch.int32)
o = 4 * B                       # ints after the doubles
lossy = raw[o:o + B]
o2 = o + ((B * 4 + 15) // 16 * 16) // 4
danger = raw[o2:o2 + 2 * B].view(B, 2)
print("lossy", lossy.tolist())
print("danger", danger.sum().item())
print("za-zb", (zl2[:, 0] - zl2[:, 1]).abs().max().item(), zl2[:4].tolist())
print("nll", nll[:6].tolist(), "U", tgl[:6], "T", inl[:6])
