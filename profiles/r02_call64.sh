mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_configs1_parity.py tests/test_gpu_module.py -q -k "split_bf16 or module or golden or configs" 2>&1 | tail -4
timeout 600 python bench.py --detail --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c64_cfg2.json 2> gpurun_out/c64_cfg2_detail.txt
python - <<'PY'
import json, re, collections
d = json.loads(open("gpurun_out/c64_cfg2.json").read().strip().split("\n")[-1])
acc = collections.OrderedDict()
for line in open("gpurun_out/c64_cfg2_detail.txt"):
    m = re.match(r"\s+(sc_\w+)\s+(\(.*?\))\s+([\d.]+) ms", line)
    if m:
        a = acc.setdefault(m.group(1), [0, 0.0]); a[0] += 1; a[1] += float(m.group(3))
print("cfg2 ms/step", d["ms_per_step"], {k: (v[0], round(v[1],3)) for k,v in acc.items() if k in ("sc_split_bf16","sc_cast","sc_colsum")})
PY
