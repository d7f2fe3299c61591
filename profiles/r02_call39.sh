mkdir -p gpurun_out
for u in 31 63 95 127 159 191 255; do echo "U in [$((u-2)),$u] (pairs per lane $(( (u+32)/32 )))"; CTC_UMAX=$u timeout 120 python profiles/ctc_time.py 2>&1 | grep lattice; done | tee gpurun_out/c38_lattice_vs_width.txt
timeout 900 python -m pytest tests/test_gpu_module.py tests/test_gpu_kernels.py tests/test_gpu_zglue_golden.py -q > gpurun_out/c39_tests.log 2>&1; tail -n 8 gpurun_out/c39_tests.log
timeout 600 python bench.py --layer-norm --detail --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c39_ln.json 2> gpurun_out/c39_ln_detail.txt
python - <<'PY'
import json, re, collections
d = json.loads(open("gpurun_out/c39_ln.json").read().strip().split("\n")[-1])
print("LN-on ms/step", d["ms_per_step"], d["clocks"], "peak_hbm_gb", d.get("peak_hbm_gb"))
acc = collections.OrderedDict()
for line in open("gpurun_out/c39_ln_detail.txt"):
    m = re.match(r"\s+(sc_\w+)\s+(\(.*?\))\s+([\d.]+) ms", line)
    if m:
        k = m.group(1)
        a = acc.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(m.group(3))
for k, (n, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:28s} n={n:3d} {t:8.3f} ms")
PY
