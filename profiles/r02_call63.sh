mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_module.py -q -k "layernorm or ln" 2>&1 | tail -4
for v in new old new old; do if [ $v = new ]; then L=""; else L=$PWD/statecatcher_b200/csrc/libsc_lnhead.so; fi; SC_B200_LIB=$L timeout 600 python bench.py --layer-norm --detail --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c63_ln.json 2> gpurun_out/c63_ln_detail.txt
python - <<PY
import json, re, collections
d = json.loads(open("gpurun_out/c63_ln.json").read().strip().split("\n")[-1])
acc = collections.OrderedDict()
for line in open("gpurun_out/c63_ln_detail.txt"):
    m = re.match(r"\s+(sc_\w+)\s+(\(.*?\))\s+([\d.]+) ms", line)
    if m:
        a = acc.setdefault(m.group(1), [0, 0.0]); a[0] += 1; a[1] += float(m.group(3))
print("$v LN-on ms/step", round(d["ms_per_step"],2), "ln_bwd", round(acc["sc_layernorm_bwd"][1],3), d["clocks"]["sm_mhz"])
PY
done
