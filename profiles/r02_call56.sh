mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rnnt.py tests/test_gpu_zglue_golden.py -q > gpurun_out/c56_tests.log 2>&1; tail -n 4 gpurun_out/c56_tests.log
timeout 900 python bench.py --workload cfg4 --detail --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/c56_cfg4.json 2> gpurun_out/c56_cfg4_detail.txt
python - <<'PY'
import json, re, collections
d = json.loads(open("gpurun_out/c56_cfg4.json").read().strip().split("\n")[-1])
print("cfg4 ms/step", d["ms_per_step"], d["clocks"])
acc = collections.OrderedDict()
for line in open("gpurun_out/c56_cfg4_detail.txt"):
    m = re.match(r"\s+(sc_\w+)\s+(\(.*?\))\s+([\d.]+) ms", line)
    if m:
        a = acc.setdefault(m.group(1), [0, 0.0]); a[0] += 1; a[1] += float(m.group(3))
for k, (n, t) in sorted(acc.items(), key=lambda kv: -kv[1][1])[:9]:
    print(f"{k:28s} n={n:4d} {t:9.3f} ms")
PY
