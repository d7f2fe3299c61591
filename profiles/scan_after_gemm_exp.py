"""Why is the forward scan slower inside the training step (0.47-0.53 ms) than alone (0.36 ms)?
Times the scan (CUDA events around it only) (a) back to back, (b) after the GEMM that produces its
input, (c) after an unrelated GEMM of the same size, (d) after an idle gap."""
import os, sys, time, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from statecatcher_b200 import ops
B, T, H = 64, 3000, 1024
M = B * T
g = torch.Generator(device='cuda').manual_seed(0)
x = torch.randn(M, H, generator=g, device='cuda').bfloat16()
w = (torch.randn(5 * H, H, generator=g, device='cuda') / 32).bfloat16()
bias = torch.zeros(5 * H, device='cuda')
G = torch.empty(M, 5 * H, device='cuda', dtype=torch.bfloat16)
G2 = torch.empty(M, 5 * H, device='cuda', dtype=torch.bfloat16)
ops.gemm_fwd(x, w, bias, out=G)
h0 = torch.zeros(B, H, device='cuda'); s0 = torch.zeros(B, H, device='cuda')

def run(pre, n=12):
    ts = []
    for i in range(n + 3):
        pre()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.scan_fwd(G, B, T, H, h0, s0, True)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]

print(f"(a) scan back to back              {run(lambda: None):.3f} ms")
print(f"(b) after the GEMM producing G     {run(lambda: ops.gemm_fwd(x, w, bias, out=G)):.3f} ms")
print(f"(c) after an unrelated GEMM        {run(lambda: ops.gemm_fwd(x, w, bias, out=G2)):.3f} ms")
def three():
    for _ in range(3):
        ops.gemm_fwd(x, w, bias, out=G2)
print(f"(d) after three unrelated GEMMs    {run(three):.3f} ms")
def gap():
    torch.cuda.synchronize(); time.sleep(0.01)
print(f"(e) after a 10 ms idle gap         {run(gap):.3f} ms")
