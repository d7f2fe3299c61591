mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c30_n1.json 2> gpurun_out/c30_n1.err
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 --steps 10 --warmup 3 --no-cpu-baseline; }
timeout 400 bash -c "$(declare -f run); run 8 29521" > gpurun_out/c30_n8.json 2> gpurun_out/c30_n8.err
timeout 400 bash -c "$(declare -f run); run 4 29522" > gpurun_out/c30_n4.json 2> gpurun_out/c30_n4.err
SC_DP_OVERLAP=1 SC_DP_GRAD=f32 timeout 400 bash -c "$(declare -f run); run 8 29523" > gpurun_out/c30_n8_overlap_f32.json 2> gpurun_out/c30_n8_overlap_f32.err
python - <<'PY'
import json
for f in ("c30_n1", "c30_n4", "c30_n8", "c30_n8_overlap_f32"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().split("\n")[-1])
        r = d["roofline_by_kernel"]
        print(f, "ms/step", round(d["ms_per_step"], 3), "frames/s", round(d["value"]), d["config"]["dp_allreduce"], {k: r[k]["ms_per_step"] for k in ("scan_fwd", "scan_bwd", "gemm", "ctc")}, d["clocks"])
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
