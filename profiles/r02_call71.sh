mkdir -p gpurun_out
# RNN-T head: one-pass joint backward (sc_joint_bwd_ws) + the [32 x 64] GEMM store boxes, configs[3] re-measured
timeout 900 python -m pytest tests/test_gpu_rnnt.py tests/test_gpu_zglue_golden.py -q -x 2>&1 | tail -4
timeout 900 python bench.py --workload cfg4 --detail --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/c71_cfg4.json 2> gpurun_out/c71_cfg4_detail.txt
python - <<'PY'
import json, re, collections
d = json.loads(open("gpurun_out/c71_cfg4.json").read().strip().split("\n")[-1])
acc = collections.OrderedDict()
for line in open("gpurun_out/c71_cfg4_detail.txt"):
    m = re.match(r"\s+(sc_\w+)\s+(\(.*?\))\s+([\d.]+) ms", line)
    if m:
        a = acc.setdefault(m.group(1)+m.group(2), [0, 0.0]); a[0] += 1; a[1] += float(m.group(3))
print("cfg4 ms/step", round(d["ms_per_step"],2), d["clocks"]["sm_mhz"])
for k,v in sorted(acc.items(), key=lambda kv:-kv[1][1])[:9]: print("  ", k, v[0], round(v[1],3))
PY
