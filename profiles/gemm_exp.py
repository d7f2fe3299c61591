"""Isolated timing of the three big cfg2 contractions (tcgen05 path)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from statecatcher_b200 import ops
M, N, K = 192000, 5120, 1024
a = torch.randn(M, K, device='cuda').bfloat16(); w = (torch.randn(N, K, device='cuda') / 32).bfloat16()
dy = torch.randn(M, N, device='cuda').bfloat16(); b = torch.randn(N, device='cuda')
y = torch.empty(M, N, device='cuda', dtype=torch.bfloat16); da = torch.empty(M, K, device='cuda', dtype=torch.bfloat16)
dw = torch.empty(N, K, device='cuda')
def t(f, n=10):
    for _ in range(3): f()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
x0 = torch.randn(M, 80, device='cuda').bfloat16(); w0 = (torch.randn(N, 80, device='cuda') / 9).bfloat16()
ms = t(lambda: ops.gemm_fwd(x0, w0, b, out=y))
print(f"fwd K=80 {ms:.3f} ms  output {M * N * 2 / ms / 1e9:.2f} TB/s")
fl = 2.0 * M * N * K
for name, f in (("fwd", lambda: ops.gemm_fwd(a, w, b, out=y)), ("dgrad", lambda: ops.gemm_dgrad(dy, w, out=da)),
                ("wgrad", lambda: ops.gemm_wgrad(dy, a, out=dw))):
    ms = t(f)
    print(f"{name:6s} {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s")
