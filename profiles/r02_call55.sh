mkdir -p gpurun_out
nvidia-smi -L | wc -l
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 --steps 10 --warmup 3 --no-cpu-baseline; }
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c55_n1.json 2> gpurun_out/c55_n1.err
timeout 400 bash -c "$(declare -f run); run 2 29531" > gpurun_out/c55_n2.json 2> gpurun_out/c55_n2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/c55_ref_n2.json 2> gpurun_out/c55_ref_n2.err; tail -c 300 gpurun_out/c55_ref_n2.json
timeout 300 python -m pytest tests/test_gpu_module.py -q -k "non_current_device" 2>&1 | tail -2
python - <<'PY'
import json
for f in ("c55_n1", "c55_n2"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().split("\n")[-1])
        r = d["roofline_by_kernel"]
        print(f, "ms/step", round(d["ms_per_step"], 3), "frames/s", round(d["value"]), d["config"]["dp_allreduce"], {k: r[k]["ms_per_step"] for k in ("scan_fwd", "scan_bwd", "gemm", "ctc")}, d["clocks"])
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
