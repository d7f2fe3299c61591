mkdir -p gpurun_out
# N=8 again with per-rank figures (own kernels / all-reduce incl. wait), NCCL_DEBUG for the algorithm, then N=1 on GPUs 0 and 7
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c75_n8.json 2> gpurun_out/c75_n8.err
grep -m3 -i "nvls\|Channel.*via\|algo" gpurun_out/c75_n8.err | head -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c75_n1.json 2> gpurun_out/c75_n1.err
CUDA_VISIBLE_DEVICES=7 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c75_n1_gpu7.json 2> gpurun_out/c75_n1_gpu7.err
python - <<'PY'
import json
for f in ("c75_n8", "c75_n1", "c75_n1_gpu7"):
    d = json.loads([l for l in open(f"gpurun_out/{f}.json").read().strip().split("\n") if l.startswith("{")][-1])
    print(f, d["n_gpus"], round(d["ms_per_step"], 2), round(d["value"]), d["clocks"]["sm_mhz"], d.get("ranks"))
PY
nvidia-smi --query-gpu=index,power.draw,power.limit,clocks.sm,temperature.gpu --format=csv
