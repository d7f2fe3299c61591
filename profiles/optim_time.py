#!/usr/bin/env python
"""Optimizer step at the cfg2 parameter set (26 tensors, 44.2 M fp32 parameters): per-tensor calls against the
multi-tensor calls, CUDA events around `step()` after warm-up, gradients re-allocated every step the way
zero_grad(set_to_none=True) leaves them.  Floor: 28 B per parameter for Adam (p, m, v read+write, g read) +
4 B for the norm = 1.41 GB -> 0.22 ms at the measured 6.55 TB/s; Lion 20 + 4 B -> 0.16 ms.
    gpurun -- 'python profiles/optim_time.py > gpurun_out/optim_time.txt'"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import statecatcher_b200 as sb  # noqa: E402
from statecatcher_b200.optim import FusedAdam, Lion  # noqa: E402

cfg = sb.LucyRNNConfig(input_dim=80, hidden_dim=1024, num_layers=6, vocab_size=1024, fused_ops=True, layer_norm=False,
                       is_training=True)


def time_opt(make, label, steps=20, warm=5):
    torch.manual_seed(0)
    model = sb.LucyRNN(cfg).cuda()
    ps = list(model.parameters())
    opt = make(ps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for it in range(warm + steps):
        for p in ps:
            p.grad = torch.randn_like(p) * 0.01
        torch.cuda.synchronize()
        e0.record()
        opt.step()
        e1.record()
        torch.cuda.synchronize()
        if it >= warm:
            tot += e0.elapsed_time(e1)
    n = sum(p.numel() for p in ps)
    print(f"{label}: {tot / steps:.4f} ms/step, {len(ps)} tensors, {n / 1e6:.1f} M parameters")


for multi in (False, True):
    time_opt(lambda ps: FusedAdam(ps, lr=1e-3, weight_decay=0.01, max_grad_norm=50.0, multi_tensor=multi),
             f"FusedAdam(AdamW, clip 50) multi_tensor={multi}")
    time_opt(lambda ps: Lion(ps, lr=1e-4, weight_decay=0.01, max_grad_norm=50.0, multi_tensor=multi),
             f"Lion(clip 50) multi_tensor={multi}")
