mkdir -p gpurun_out
# fused backward scan: 8-row stages x3 (default) against 16-row stages x2 (libsc_b16.so, -DSC_SCAN_BWD_ROWS=16)
SC_B200_LIB=$PWD/statecatcher_b200/csrc/libsc_b16.so timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_configs1_parity.py tests/test_gpu_module.py -q -k "scan or golden or configs or module" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_configs1_parity.py -q -k "scan or configs" 2>&1 | tail -3
for i in 1 2; do for lib in libstatecatcher_b200.so libsc_b16.so; do
SC_B200_LIB=$PWD/statecatcher_b200/csrc/$lib timeout 600 python bench.py --detail --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c67.json 2> gpurun_out/c67_detail.txt
python - $lib <<'PY'
import json, re, collections, sys
d = json.loads(open("gpurun_out/c67.json").read().strip().split("\n")[-1])
acc = collections.OrderedDict()
for line in open("gpurun_out/c67_detail.txt"):
    m = re.match(r"\s+(sc_\w+)\s+(\(.*?\))\s+([\d.]+) ms", line)
    if m:
        a = acc.setdefault(m.group(1), [0, 0.0]); a[0] += 1; a[1] += float(m.group(3))
print(sys.argv[1], "ms/step", round(d["ms_per_step"],2), {k: round(v[1],3) for k,v in acc.items() if "scan" in k}, d["clocks"]["sm_mhz"])
PY
done; done
