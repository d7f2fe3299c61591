mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_module.py tests/test_gpu_kernels.py -q -k "layernorm or ln or golden or module" > gpurun_out/c57_tests.log 2>&1; tail -n 4 gpurun_out/c57_tests.log
for v in 1 0 1 0; do SC_LN_BWD_STAGE=$v timeout 600 python bench.py --layer-norm --detail --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c57_ln.json 2> gpurun_out/c57_ln_detail.txt
python - <<PY
import json, re, collections
d = json.loads(open("gpurun_out/c57_ln.json").read().strip().split("\n")[-1])
acc = collections.OrderedDict()
for line in open("gpurun_out/c57_ln_detail.txt"):
    m = re.match(r"\s+(sc_\w+)\s+(\(.*?\))\s+([\d.]+) ms", line)
    if m:
        a = acc.setdefault(m.group(1), [0, 0.0]); a[0] += 1; a[1] += float(m.group(3))
print("stage=$v LN-on ms/step", round(d["ms_per_step"],2), "ln_bwd", round(acc["sc_layernorm_bwd"][1],3), "ln_fwd", round(acc["sc_layernorm_fwd"][1],3), d["clocks"]["sm_mhz"])
PY
done
