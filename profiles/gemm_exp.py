import torch, sys, os
sys.path.insert(0, '/root/repo')
from statecatcher_b200 import ops
M=192000
for K in (80, 1024):
    a=torch.randn(M,K,device='cuda').bfloat16(); w=torch.randn(5120,K,device='cuda').bfloat16(); b=torch.randn(5120,device='cuda')
    out=torch.empty(M,5120,device='cuda',dtype=torch.bfloat16)
    for _ in range(3): ops.gemm_fwd(a,w,b,out=out)
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.gemm_fwd(a,w,b,out=out)
    e1.record(); torch.cuda.synchronize()
    print(os.environ.get('SC_GEMM_DBG','0'), K, e0.elapsed_time(e1)/10)
