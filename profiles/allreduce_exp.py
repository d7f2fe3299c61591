"""Gradient all-reduce alone: the cfg2 payload (44.3 M gradients) as bf16 / fp32, ReduceOp.AVG against SUM.
torchrun --nproc-per-node N profiles/allreduce_exp.py"""
import os, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 44_300_000
for dtype in (torch.bfloat16, torch.float32):
    buf = torch.randn(n, device="cuda").to(dtype)
    for op, name in ((dist.ReduceOp.AVG, "AVG"), (dist.ReduceOp.SUM, "SUM")):
        for _ in range(5):
            dist.all_reduce(buf, op=op)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            dist.all_reduce(buf, op=op)
            if op == dist.ReduceOp.SUM:
                buf.mul_(1.0 / dist.get_world_size())
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        if rank == 0:
            print(f"world {dist.get_world_size()} {str(dtype):16s} {name}  {ms:.3f} ms per all-reduce (+ scale for SUM)  {buf.numel() * buf.element_size() / 1e6:.0f} MB", flush=True)
dist.destroy_process_group()
