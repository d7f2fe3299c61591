mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 300 python -m pytest tests/test_gpu_module.py -q -k "non_current_device" > gpurun_out/c21_twodev.log 2>&1; tail -n 3 gpurun_out/c21_twodev.log
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline; }
timeout 600 bash -c "$(declare -f run); run 29511" > gpurun_out/c21_n2_default.json 2> gpurun_out/c21_n2_default.err
SC_DP_OVERLAP=1 SC_DP_GRAD=f32 timeout 600 bash -c "$(declare -f run); run 29512" > gpurun_out/c21_n2_overlap_f32.json 2> gpurun_out/c21_n2_overlap_f32.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c21_n1.json 2> gpurun_out/c21_n1.err
python - <<'PY'
import json
for f in ("c21_n1", "c21_n2_default", "c21_n2_overlap_f32"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().split("\n")[-1])
        r = d["roofline_by_kernel"]
        print(f, "ms/step", round(d["ms_per_step"], 3), "frames/s", round(d["value"]), d["config"]["dp_allreduce"], {k: r[k]["ms_per_step"] for k in ("scan_fwd", "scan_bwd", "gemm")})
    except Exception as e:
        print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
