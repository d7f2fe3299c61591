mkdir -p gpurun_out
# LayerNorm forward: persistent pipelined kernel (default) against one row per warp (SC_LN_FWD_PIPE=0)
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_module.py -q -k "layernorm or ln or golden or module" 2>&1 | tail -3
for i in 1 2; do for pipe in 1 0; do
SC_LN_FWD_PIPE=$pipe timeout 600 python bench.py --layer-norm --detail --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c69_ln.json 2> gpurun_out/c69_ln_detail.txt
python - $pipe <<'PY'
import json, re, collections, sys
d = json.loads(open("gpurun_out/c69_ln.json").read().strip().split("\n")[-1])
acc = collections.OrderedDict()
for line in open("gpurun_out/c69_ln_detail.txt"):
    m = re.match(r"\s+(sc_\w+)\s+(\(.*?\))\s+([\d.]+) ms", line)
    if m:
        a = acc.setdefault(m.group(1), [0, 0.0]); a[0] += 1; a[1] += float(m.group(3))
print("pipe", sys.argv[1], "LN ms/step", round(d["ms_per_step"],2), {k: round(v[1],3) for k,v in acc.items() if "layernorm" in k}, d["clocks"]["sm_mhz"])
PY
done; done
