import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from statecatcher_b200 import ops
M, N, K = 192000, 5120, 1024
dy = torch.randn(M, N, device='cuda').bfloat16(); a = torch.randn(M, K, device='cuda').bfloat16()
out = torch.empty(N, K, device='cuda')
for _ in range(3): ops.gemm_wgrad(dy, a, out=out)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.gemm_wgrad(dy, a, out=out)
e1.record(); torch.cuda.synchronize()
ref = (dy[:4096].float().T @ a[:4096].float())
chk = torch.empty(N, K, device='cuda'); ops.gemm_wgrad(dy[:4096], a[:4096], out=chk)
print(os.environ.get('SC_WGRAD_SPLITS', 'auto'), e0.elapsed_time(e1) / 10, "ms  maxrel", ((chk - ref).abs().max() / ref.abs().max()).item())
