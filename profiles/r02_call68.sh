mkdir -p gpurun_out
# GEMM epilogue: [32 x 64] store boxes (128-byte rows) against the [32 x 32] boxes of libsc_old.so
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_kernels.py tests/test_gpu_configs1_parity.py tests/test_gpu_module.py -q -x 2>&1 | tail -4
for i in 1 2; do for lib in libstatecatcher_b200.so libsc_old.so; do
SC_B200_LIB=$PWD/statecatcher_b200/csrc/$lib timeout 600 python bench.py --detail --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c68.json 2> gpurun_out/c68_detail.txt
python - $lib <<'PY'
import json, re, collections, sys
d = json.loads(open("gpurun_out/c68.json").read().strip().split("\n")[-1])
acc = collections.OrderedDict()
for line in open("gpurun_out/c68_detail.txt"):
    m = re.match(r"\s+(sc_\w+)\s+(\(.*?\))\s+([\d.]+) ms", line)
    if m:
        a = acc.setdefault(m.group(1)+m.group(2), [0, 0.0]); a[0] += 1; a[1] += float(m.group(3))
print(sys.argv[1], "ms/step", round(d["ms_per_step"],2), {k: round(v[1],3) for k,v in acc.items() if "gemm" in k and "192000" in k}, d["clocks"]["sm_mhz"])
PY
done; done
