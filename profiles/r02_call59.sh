mkdir -p gpurun_out
SC_B200_LIB=$PWD/statecatcher_b200/csrc/libsc_deep.so timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "scan" > gpurun_out/c59_tests.log 2>&1; tail -n 3 gpurun_out/c59_tests.log
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline_by_kernel']
print('RESULT', round(d['ms_per_step'],2), 'scan_fwd', round(r['scan_fwd']['ms_per_step'],3), r['scan_fwd']['frac'], 'scan_bwd', round(r['scan_bwd']['ms_per_step'],3), r['scan_bwd']['frac'], 'gemm', round(r['gemm']['ms_per_step'],2), d['clocks']['sm_mhz'])"; }
lnrun() { python bench.py --layer-norm --detail --steps 5 --warmup 3 --no-cpu-baseline 2> gpurun_out/c59_ln_detail.txt | tail -1 > gpurun_out/c59_ln.json; python - <<PY
import json, re, collections
d = json.loads(open("gpurun_out/c59_ln.json").read().strip().split("\n")[-1])
acc = collections.OrderedDict()
for line in open("gpurun_out/c59_ln_detail.txt"):
    m = re.match(r"\s+(sc_\w+)\s+(\(.*?\))\s+([\d.]+) ms", line)
    if m:
        a = acc.setdefault(m.group(1), [0, 0.0]); a[0] += 1; a[1] += float(m.group(3))
print("LN ms/step", round(d["ms_per_step"],2), {k: round(v[1],3) for k,v in acc.items() if "scan" in k})
PY
}
echo base; run; lnrun
echo deep; export SC_B200_LIB=$PWD/statecatcher_b200/csrc/libsc_deep.so; run; lnrun; unset SC_B200_LIB
echo base; run
echo deep; export SC_B200_LIB=$PWD/statecatcher_b200/csrc/libsc_deep.so; run; unset SC_B200_LIB
