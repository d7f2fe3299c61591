#!/usr/bin/env python
"""What slows the CTC recursions when another kernel shares the device?  The lattice call (cfg2 shape) on one stream,
timed by events on that stream, while a second stream runs (a) a DRAM copy, (b) an issue-heavy kernel on an
L2-resident tensor, (c) the emission pass, (d) the gradient pass — each long enough to cover the recursions."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402
from statecatcher_b200._lib import call, dt, load, ptr  # noqa: E402

W = bench.WORKLOADS["cfg2"]
_, tok, inl, tgl = bench.synth_batch(W, 1234)
B, T, V = W["B"], W["T"], W["V"]
g = torch.Generator(device="cuda").manual_seed(0)
x = (torch.randn(B, T, V, generator=g, device="cuda") * 2).bfloat16()
tok = tok.cuda()
U = int(max(tgl))
il, tl = torch.tensor(inl).cuda(), torch.tensor(tgl).cuda()
S = (2 * U + 1 + 7) & ~7
f32 = dict(dtype=torch.float32, device="cuda")


def bufs():
    return dict(lse=torch.zeros(B, T, **f32), lplat=torch.zeros(B, T, load().sc_ctc_lplat_pitch(U), **f32), csh=torch.zeros(B, T, **f32),
                alpha=torch.zeros(B, T, S, **f32), beta=torch.zeros(B, T, S, **f32), nll=torch.zeros(B, **f32),
                loss=torch.zeros((), **f32), dx=torch.empty_like(x),
                ws=torch.zeros(load().sc_ctc_workspace_bytes(B, T, U) // 8 + 1, dtype=torch.float64, device="cuda"))


b1, b2 = bufs(), bufs()
one = torch.ones((), **f32)
sa, sb_ = torch.cuda.Stream(priority=-1), torch.cuda.Stream()


def emis(b, st):
    call("sc_ctc_emissions", ptr(x), x.stride(0), x.stride(1), dt(x), ptr(tok), tok.stride(0), ptr(il), ptr(tl),
         B, T, V, U, 0, ptr(b["lse"]), ptr(b["lplat"]), ptr(b["csh"]), st.cuda_stream)


def lat(b, st):
    call("sc_ctc_lattice", ptr(b["lplat"]), ptr(b["csh"]), ptr(tok), tok.stride(0), ptr(il), ptr(tl), B, T, U, 0,
         ptr(b["alpha"]), ptr(b["beta"]), ptr(b["nll"]), ptr(b["loss"]), 1, ptr(b["ws"]), st.cuda_stream)


def grad(b, st):
    call("sc_ctc_bwd", ptr(x), x.stride(0), x.stride(1), dt(x), ptr(tok), tok.stride(0), ptr(il), ptr(tl),
         B, T, V, U, 0, ptr(b["lse"]), ptr(b["alpha"]), ptr(b["beta"]), ptr(b["nll"]), ptr(one), 1,
         ptr(b["dx"]), b["dx"].stride(0), b["dx"].stride(1), dt(b["dx"]), ptr(b["ws"]), st.cuda_stream)


cur = torch.cuda.current_stream()
for b in (b1, b2):
    emis(b, cur); lat(b, cur); grad(b, cur)
torch.cuda.synchronize()
big_a = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
big_b = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
small = torch.randn(4 << 20, device="cuda")


def other(kind):
    with torch.cuda.stream(sb_):
        if kind == "copy":
            for _ in range(2):
                big_b.copy_(big_a)
        elif kind == "math":
            y = small
            for _ in range(40):
                y = torch.erfinv(torch.tanh(y) * 0.5)
        elif kind == "emissions":
            for _ in range(4):
                emis(b2, sb_)
        elif kind == "grad":
            for _ in range(2):
                grad(b2, sb_)


for kind in (None, "copy", "math", "emissions", "grad"):
    ts, to = [], []
    for i in range(5):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o0.record(sb_)
        if kind:
            other(kind)
        o1.record(sb_)
        e0.record(sa)
        lat(b1, sa)
        e1.record(sa)
        torch.cuda.synchronize()
        if i >= 1:
            ts.append(e0.elapsed_time(e1)); to.append(o0.elapsed_time(o1))
    print("recursions %.4f ms (min %.4f) next to %-10s (which took %.4f ms)" % (sum(ts) / len(ts), min(ts), kind, sum(to) / len(to)))
