mkdir -p gpurun_out
for v in 3 4 6 3 4 6; do if [ $v = 3 ]; then L=""; else L=$PWD/statecatcher_b200/csrc/libsc_ring$v.so; fi; SC_B200_LIB=$L timeout 600 python bench.py --layer-norm --detail --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c58_ln.json 2> gpurun_out/c58_ln_detail.txt
python - <<PY
import json, re, collections
d = json.loads(open("gpurun_out/c58_ln.json").read().strip().split("\n")[-1])
acc = collections.OrderedDict()
for line in open("gpurun_out/c58_ln_detail.txt"):
    m = re.match(r"\s+(sc_\w+)\s+(\(.*?\))\s+([\d.]+) ms", line)
    if m:
        a = acc.setdefault(m.group(1), [0, 0.0]); a[0] += 1; a[1] += float(m.group(3))
print("ring=$v LN-on ms/step", round(d["ms_per_step"],2), "ln_bwd", round(acc["sc_layernorm_bwd"][1],3), d["clocks"]["sm_mhz"])
PY
done
