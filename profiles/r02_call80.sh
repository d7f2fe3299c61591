mkdir -p gpurun_out
timeout 600 python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c80_cfg4.json 2> gpurun_out/c80_cfg4.err
timeout 300 python bench.py --workload cfg1 --graph --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c80_cfg1_graph.json 2>/dev/null
python - <<'PY'
import json
for f in ("c80_cfg4", "c80_cfg1_graph"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().split("\n")[-1])
    print(f, round(d["ms_per_step"], 3), round(d["value"]), d["clocks"]["sm_mhz"], d["roofline"]["kernel"], round(d["roofline"]["frac"], 3))
PY
