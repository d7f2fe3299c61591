import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from statecatcher_b200 import ops
B, T, H = 64, 3000, 1024
g = torch.Generator(device='cuda').manual_seed(0)
G = torch.randn(B * T, 5 * H, generator=g, device='cuda').bfloat16()
h0 = torch.zeros(B, H, device='cuda'); s0 = torch.zeros(B, H, device='cuda')
go = torch.randn(B * T, H, generator=g, device='cuda').bfloat16()
def timeit(f, n=10):
    for _ in range(3): f()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
Hout, hT, _, ck = ops.scan_fwd(G, B, T, H, h0, s0, True)
tf = timeit(lambda: ops.scan_fwd(G, B, T, H, h0, s0, True))
tb = timeit(lambda: ops.scan_bwd(G, Hout, h0, s0, ck, go, B, T, H, True))
fb = 6 * H * 2 * B * T; bb = 12 * H * 2 * B * T
print(f"VEC={os.environ.get('SC_SCAN_VEC','2')} GENERIC={os.environ.get('SC_SCAN_GENERIC','0')} fwd {tf:.3f} ms {fb/tf/1e6:.0f} GB/s  bwd {tb:.3f} ms {bb/tb/1e6:.0f} GB/s")
