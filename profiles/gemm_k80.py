"""Layer-0 projection alone (192000 x 5120 x 80, output-bound) — for ncu."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from statecatcher_b200 import ops
M, N, K = 192000, 5120, 80
x = torch.randn(M, K, device='cuda').bfloat16(); w = (torch.randn(N, K, device='cuda') / 9).bfloat16()
b = torch.randn(N, device='cuda'); y = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
for _ in range(3): ops.gemm_fwd(x, w, b, out=y)
torch.cuda.synchronize(); print("ok")
