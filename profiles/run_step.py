#!/usr/bin/env python
"""Minimal driver for profiling: N training steps (fwd + CTC + bwd, carried state) of the
bench.py workload, nothing else.  Used under ncu (see profiles/README.md)."""
import argparse
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import bench  # noqa: E402
import statecatcher_b200 as sb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--layer-norm", action="store_true")
a = ap.parse_args()
W = bench.WORKLOADS[a.workload]
cd = torch.bfloat16 if W["dtype"] == "bf16" else torch.float32
cfg = sb.LucyRNNConfig(input_dim=W["F"], hidden_dim=W["H"], num_layers=W["L"], vocab_size=W["V"],
                       is_training=True, fused_ops=True, layer_norm=a.layer_norm)
model = sb.LucyRNN(cfg, compute_dtype=cd)
bench.init_reference_like(model, 1234)
model = model.cuda()
x, tok, inl, tgl = bench.synth_batch(W, 1234)
x, tok = x.cuda(), tok.cuda()
inl, tgl = torch.tensor(inl).cuda(), torch.tensor(tgl).cuda()
state = None
for i in range(a.steps):
    st = sb.detach_states(state) if state else None
    model.zero_grad(set_to_none=True)
    logits, state = model(x, st) if st else model(x)
    loss = sb.ctc_loss_from_logits(logits, tok, inl, tgl, zero_infinity=True)
    loss.backward()
torch.cuda.synchronize()
print("ok", loss.item())
