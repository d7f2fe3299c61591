mkdir -p gpurun_out
timeout 300 python profiles/ctc_head_exp.py > gpurun_out/c27_head_exp.txt 2>&1; cat gpurun_out/c27_head_exp.txt
for p in 4 6 8; do SC_CTC_PHASES=$p timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c27_n1_p$p.json 2> gpurun_out/c27_n1_p$p.err; python - <<PY
import json
d = json.loads(open("gpurun_out/c27_n1_p$p.json").read().strip().split("\n")[-1])
print("phases $p ms/step", d["ms_per_step"], "ctc", d["roofline_by_kernel"]["ctc"], d["clocks"])
PY
done
