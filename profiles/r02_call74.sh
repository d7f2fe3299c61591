mkdir -p gpurun_out
# final code on one 8-GPU box: N=8, then N=1 (same box)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c74_n8.json 2> gpurun_out/c74_n8.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c74_n1.json 2> gpurun_out/c74_n1.err
python - <<'PY'
import json
for f in ("c74_n8", "c74_n1"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().split("\n")[-1])
    print(f, d["n_gpus"], round(d["ms_per_step"], 2), round(d["value"]), d["clocks"]["sm_mhz"], d["config"].get("dp_allreduce"))
PY
