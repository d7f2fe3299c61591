mkdir -p gpurun_out
# N=2 sanity of the final code (torchrun path, flat bf16 all-reduce), then N=1 on the same box
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c73_n2.json 2> gpurun_out/c73_n2.err; tail -c 600 gpurun_out/c73_n2.json | head -c 600; echo
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c73_n1.json 2> gpurun_out/c73_n1.err
python - <<'PY'
import json
for f in ("c73_n2", "c73_n1"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().split("\n")[-1])
    print(f, d["n_gpus"], round(d["ms_per_step"], 2), round(d["value"]), d["clocks"]["sm_mhz"], d["config"].get("dp_allreduce"))
PY
timeout 600 python -m pytest tests -m gpu -q -k "device_guard or non_current or two_gpu or 2gpu" 2>&1 | tail -3
