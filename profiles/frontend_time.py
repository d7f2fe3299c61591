#!/usr/bin/env python
"""Frontend kernel alone at the configs[1] shape: 64 streams x 30 s at 16 kHz -> (64, 2998, 80) MFCC.
Algorithmic bytes per frame: 160 samples x 4 B in + 80 x 4 B out = 960 B (csrc/sc_frontend.cu)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import statecatcher_b200 as sb  # noqa: E402

B, S = 64, 480000
wav = torch.randn(B, S, device="cuda") * 0.1
for kind in ("mfcc", "mel"):
    fe = sb.make_frontend(kind, 16000)[0].cuda()
    for _ in range(3):
        out = fe.features(wav)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        out = fe.features(wav)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    frames = B * out.shape[1]
    print(f"{kind}: {ms:.3f} ms per batch, {frames / ms * 1e3 / 1e6:.1f} M frames/s, "
          f"{frames * 960 / ms / 1e6:.0f} GB/s algorithmic, {frames * 2 * 40.6e3 / ms / 1e9:.1f} TFLOP/s fp32 (folded DFT only)")
mask = torch.ones(B, S, dtype=torch.bool, device="cuda")
for _ in range(3):
    fm, lens = sb.frame_mask_and_lens(mask, 2998)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    fm, lens = sb.frame_mask_and_lens(mask, 2998)
e1.record()
torch.cuda.synchronize()
print(f"frame_mask+in_lens: {e0.elapsed_time(e1) / 10:.3f} ms per batch (includes the lens D2H)")
