mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_module.py tests/test_gpu_kernels.py -q -k "split or ln or golden or module" > gpurun_out/c43_tests.log 2>&1; tail -n 5 gpurun_out/c43_tests.log
for i in 1 2; do timeout 600 python bench.py --layer-norm --detail --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c43_ln.json 2> gpurun_out/c43_ln_detail.txt
python - <<'PY'
import json, re, collections
d = json.loads(open("gpurun_out/c43_ln.json").read().strip().split("\n")[-1])
print("LN-on ms/step", d["ms_per_step"], d["clocks"], "peak_hbm_gb", d.get("peak_hbm_gb"))
acc = collections.OrderedDict()
for line in open("gpurun_out/c43_ln_detail.txt"):
    m = re.match(r"\s+(sc_\w+)\s+(\(.*?\))\s+([\d.]+) ms", line)
    if m:
        k = m.group(1)
        a = acc.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(m.group(3))
for k, (n, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    if "scan" in k or "layernorm" in k: print(f"{k:28s} n={n:3d} {t:8.3f} ms")
PY
done
