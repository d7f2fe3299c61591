mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c19_gpu_suite.log 2>&1
timeout 300 python bench.py --workload cfg1 --steps 10 --warmup 3 --detail --no-cpu-baseline > gpurun_out/c19_bench_cfg1.json 2> gpurun_out/c19_bench_cfg1_detail.txt
timeout 300 python bench.py --workload cfg1 --steps 10 --warmup 3 --graph > gpurun_out/c19_bench_cfg1_graph.json 2> gpurun_out/c19_bench_cfg1_graph.err
tail -n 8 gpurun_out/c19_gpu_suite.log; tail -3 gpurun_out/c19_bench_cfg1_graph.err
python - <<'PY'
import json
for f in ("c19_bench_cfg1", "c19_bench_cfg1_graph"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().split("\n")[-1])
        print(f, "ms/step", round(d["ms_per_step"], 3), "frames/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"], d.get("cpu_baseline"))
    except Exception as e:
        print(f, "failed", e)
PY
