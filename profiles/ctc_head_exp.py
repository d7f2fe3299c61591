#!/usr/bin/env python
"""sc_ctc_head at the cfg2 shape: the passes one after the other vs the overlapped head for several phase counts,
with and without the high-priority recursion stream (CUDA events on the calling stream around the C-ABI call)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402
from statecatcher_b200 import ctc  # noqa: E402
from statecatcher_b200._lib import call, dt, load, ptr, stream  # noqa: E402

W = bench.WORKLOADS[os.environ.get("WORKLOAD", "cfg2")]
_, tok, inl, tgl = bench.synth_batch(W, 1234)
B, T, V = W["B"], W["T"], W["V"]
g = torch.Generator(device="cuda").manual_seed(0)
x = (torch.randn(B, T, V, generator=g, device="cuda") * 2).bfloat16()
tok = tok.cuda()
U = int(max(tgl))
il, tl = torch.tensor(inl).cuda(), torch.tensor(tgl).cuda()
S = (2 * U + 1 + 7) & ~7
f32 = dict(dtype=torch.float32, device="cuda")
lse, lplat, csh = torch.zeros(B, T, **f32), torch.zeros(B, T, load().sc_ctc_lplat_pitch(U), **f32), torch.zeros(B, T, **f32)
alpha, beta = torch.zeros(B, T, S, **f32), torch.zeros(B, T, S, **f32)
nll, loss, one = torch.zeros(B, **f32), torch.zeros((), **f32), torch.ones((), **f32)
ws = torch.zeros(load().sc_ctc_workspace_bytes(B, T, U) // 8 + 1, dtype=torch.float64, device="cuda")
dx = torch.empty_like(x)
sl, se, sg = ctc._head_streams(x.device)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def seq():
    call("sc_ctc_fwd", ptr(x), x.stride(0), x.stride(1), dt(x), ptr(tok), tok.stride(0), ptr(il), ptr(tl),
         B, T, V, U, 0, ptr(lse), ptr(lplat), ptr(csh), ptr(alpha), ptr(beta), ptr(nll), ptr(loss), 1, ptr(ws), stream())
    call("sc_ctc_bwd", ptr(x), x.stride(0), x.stride(1), dt(x), ptr(tok), tok.stride(0), ptr(il), ptr(tl),
         B, T, V, U, 0, ptr(lse), ptr(alpha), ptr(beta), ptr(nll), ptr(one), 1,
         ptr(dx), dx.stride(0), dx.stride(1), dt(dx), ptr(ws), stream())


def head(P, prio=True, side=True):
    call("sc_ctc_head", ptr(x), x.stride(0), x.stride(1), dt(x), ptr(tok), tok.stride(0), ptr(il), ptr(tl),
         B, T, V, U, 0, ptr(lse), ptr(lplat), ptr(csh), ptr(alpha), ptr(beta), ptr(nll), ptr(loss), 1, ptr(ws),
         ptr(dx), dx.stride(0), dx.stride(1), dt(dx), P, stream(), sl if prio else 0, se if side else 0, sg if side else 0)


def timeit(fn, n=6):
    ts = []
    for i in range(n + 2):
        flush.zero_()
        torch.cuda._sleep(3_000_000)          # ~1.5 ms of device time: the host gets ahead, as it is inside a training step
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    return sum(ts) / len(ts), min(ts)


print("shape", B, T, V, "Umax", U)
print("one after the other      mean %.4f min %.4f ms" % timeit(seq))
ref = dx.clone()
for P in (2, 3, 4, 5, 6, 7, 8, 10, 12, 16):
    a = timeit(lambda: head(P, True))
    same = torch.equal(dx.view(torch.int16), ref.view(torch.int16))
    b = timeit(lambda: head(P, False))
    print("head phases %2d (-> %2d)  priority stream: mean %.4f min %.4f   recursions on the caller's stream: mean %.4f min %.4f   same=%s"
          % (P, load().sc_ctc_head_phases(T, U, P), a[0], a[1], b[0], b[1], same))
print("loss", loss.item())
