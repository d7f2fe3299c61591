mkdir -p gpurun_out
timeout 120 python profiles/ctc_time.py > gpurun_out/c17_ctc_time.txt 2>&1
for v in 2 1; do
  SC_SCAN_VEC=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c17_bench_vec$v.json 2> gpurun_out/c17_bench_vec$v.err
done
cat gpurun_out/c17_ctc_time.txt
python - <<'PY'
import json
for v in (2, 1):
    d = json.loads(open(f"gpurun_out/c17_bench_vec{v}.json").read().strip().split("\n")[-1])
    r = d["roofline_by_kernel"]
    print("VEC", v, "ms/step", round(d["ms_per_step"], 3), {k: (r[k]["ms_per_step"], r[k]["frac"]) for k in ("scan_fwd", "scan_bwd", "ctc", "gemm")}, d["clocks"])
PY
