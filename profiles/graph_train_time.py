#!/usr/bin/env python
"""BASELINE configs[0] (2 x 256, V=1024, 8 streams x 1000 frames, fp32): training step eager against
GraphedTrainStep, CUDA events over 20 steps after warm-up, carried state.  NOT yet run (round 1's GPU budget
was spent when it was written).
    gpurun -- 'python profiles/graph_train_time.py > gpurun_out/graph_train_time.txt'"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import statecatcher_b200 as sb  # noqa: E402

B, T, F, H, L, V, U = 8, 1000, 80, 256, 2, 1024, 50
cfg = sb.LucyRNNConfig(input_dim=F, hidden_dim=H, num_layers=L, vocab_size=V, fused_ops=True, layer_norm=False,
                       is_training=True)
torch.manual_seed(0)
model = sb.LucyRNN(cfg).cuda()
with torch.no_grad():
    model.output_proj.weight.normal_(0, 0.02)
x = torch.randn(B, T, F, device="cuda")
tok = torch.randint(1, V, (B, U), device="cuda")
inl = torch.full((B,), T, device="cuda", dtype=torch.int64)
tgl = torch.randint(25, U + 1, (B,), device="cuda", dtype=torch.int64)
crit = sb.CTCLoss(blank=0, zero_infinity=True)


def timed(fn, steps=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


state = [None]


def eager():
    model.zero_grad(set_to_none=True)
    st = sb.detach_states(state[0]) if state[0] else None
    logits, state[0] = model(x, st) if st else model(x)
    crit(logits.transpose(0, 1), tok, inl, tgl).backward()


ms = timed(eager)
print(f"eager   : {ms:.3f} ms/step, {B * T / ms * 1e3:.0f} frames/s")
runner = sb.GraphedTrainStep(model, batch=B, frames=T, feat_dim=F, max_labels=U)
ms = timed(lambda: runner.step(x, tok, inl, tgl))
print(f"graphed : {ms:.3f} ms/step, {B * T / ms * 1e3:.0f} frames/s   loss {runner.loss.item():.4f}")
