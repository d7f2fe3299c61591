mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c18_gpu_suite.log 2>&1
timeout 300 python bench.py --workload cfg1 --steps 10 --warmup 3 --detail --no-cpu-baseline > gpurun_out/c18_bench_cfg1.json 2> gpurun_out/c18_bench_cfg1_detail.txt
SC_F32_GEMM=simt timeout 300 python bench.py --workload cfg1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c18_bench_cfg1_simt.json 2>/dev/null
tail -n 15 gpurun_out/c18_gpu_suite.log
python - <<'PY'
import json
for f in ("c18_bench_cfg1", "c18_bench_cfg1_simt"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().split("\n")[-1])
    print(f, "ms/step", round(d["ms_per_step"], 3), "frames/s", round(d["value"]), "launches", d["gpu_launches"])
PY
