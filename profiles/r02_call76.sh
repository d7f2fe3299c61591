mkdir -p gpurun_out
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/c76_nccl_%h_%p.log timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 profiles/allreduce_exp.py 2>&1 | grep -v "^\*\|OMP_NUM" | tee gpurun_out/c76_allreduce.txt
cat gpurun_out/c76_nccl_*.log | grep -i "nvls\|via\|Connected\|channels\|algo" | sort | uniq -c | sort -rn | head -12
rm -f gpurun_out/c76_nccl_*.log
