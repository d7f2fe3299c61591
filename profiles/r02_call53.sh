mkdir -p gpurun_out
timeout 600 python bench.py --detail --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c53_cfg2.json 2> gpurun_out/c53_cfg2_detail.txt
python - <<'PY'
import json, re, collections
d = json.loads(open("gpurun_out/c53_cfg2.json").read().strip().split("\n")[-1])
print("cfg2 ms/step", d["ms_per_step"], d["clocks"])
acc = collections.OrderedDict()
for line in open("gpurun_out/c53_cfg2_detail.txt"):
    m = re.match(r"\s+(sc_\w+)\s+(\(.*?\))\s+([\d.]+) ms\s+([\d.]+)", line)
    if m:
        k = m.group(1) + " " + m.group(2)
        a = acc.setdefault(k, [0, 0.0, []]); a[0] += 1; a[1] += float(m.group(3)); a[2].append(float(m.group(4)))
tot = sum(v[1] for v in acc.values())
print("sum of calls", tot)
for k, (n, t, r) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:48s} n={n:3d} {t:8.3f} ms  rate min/max {min(r):8.1f} {max(r):8.1f}")
PY
