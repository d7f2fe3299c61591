mkdir -p gpurun_out
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING NCCL_DEBUG_FILE=gpurun_out/c77_nccl_%h_%p.log timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 profiles/allreduce_exp.py 2>&1 | grep -v "^\*\|OMP_NUM" | tee gpurun_out/c77_allreduce.txt
cat gpurun_out/c77_nccl_*.log | grep -i "nvls\|Connected\|AllReduce:" | sed 's/^.*NCCL INFO //' | sort | uniq -c | sort -rn | head -12
rm -f gpurun_out/c77_nccl_*.log
