mkdir -p gpurun_out
timeout 900 python bench.py --workload cfg4 --detail --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/c49_cfg4.json 2> gpurun_out/c49_cfg4_detail.txt
python - <<'PY'
import json, re, collections
d = json.loads(open("gpurun_out/c49_cfg4.json").read().strip().split("\n")[-1])
print("cfg4 ms/step", d["ms_per_step"], d["clocks"], "peak_hbm_gb", d.get("peak_hbm_gb"))
acc = collections.OrderedDict(); shapes = collections.Counter()
for line in open("gpurun_out/c49_cfg4_detail.txt"):
    m = re.match(r"\s+(sc_\w+)\s+(\(.*?\))\s+([\d.]+) ms\s+([\d.]+)", line)
    if m:
        k = m.group(1) + " " + m.group(2)
        a = acc.setdefault(k, [0, 0.0, 0.0]); a[0] += 1; a[1] += float(m.group(3)); a[2] = float(m.group(4))
for k, (n, t, r) in sorted(acc.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f"{k:60s} n={n:4d} {t:9.3f} ms  last rate {r}")
PY
