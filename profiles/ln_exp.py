"""LayerNorm forward / backward alone at the cfg2 shape ([192000 x 1024] bf16), contiguous and as a column block of the
gate tensor (row stride 5H), next to a device copy of the same bytes.  Used plain and under ncu."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from statecatcher_b200 import ops
M, H = 192000, 1024
g = torch.Generator(device='cuda').manual_seed(0)
G = torch.randn(M, 5 * H, generator=g, device='cuda').bfloat16()
x = torch.randn(M, H, generator=g, device='cuda').bfloat16()
dy = torch.randn(M, H, generator=g, device='cuda').bfloat16()
w = torch.randn(H, device='cuda'); b = torch.randn(H, device='cuda')
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
def timeit(f, n=10):
    for _ in range(3): f()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]
y, mean, rstd = ops.layernorm_fwd(x, w, b)
dx = torch.empty_like(x)
t = timeit(lambda: ops.layernorm_fwd(x, w, b, out=y)); print(f"ln fwd contiguous     {t:.3f} ms {2*M*H*2/t/1e6:.0f} GB/s")
t = timeit(lambda: ops.layernorm_fwd(G[:, :H], w, b, out=y)); print(f"ln fwd block of G     {t:.3f} ms {2*M*H*2/t/1e6:.0f} GB/s")
t = timeit(lambda: ops.layernorm_bwd(dy, x, w, mean, rstd, dx=dx)); print(f"ln bwd contiguous     {t:.3f} ms {3*M*H*2/t/1e6:.0f} GB/s")
t = timeit(lambda: y.copy_(x)); print(f"copy                  {t:.3f} ms {2*M*H*2/t/1e6:.0f} GB/s")
