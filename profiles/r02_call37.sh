mkdir -p gpurun_out
SC_B200_LIB=$PWD/statecatcher_b200/csrc/libstatecatcher_b200_cb128.so timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "scan" > gpurun_out/c37_cb128_tests.log 2>&1; tail -n 3 gpurun_out/c37_cb128_tests.log
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline_by_kernel']
print('RESULT', round(d['ms_per_step'],2), 'scan_fwd', round(r['scan_fwd']['ms_per_step'],3), r['scan_fwd']['frac'], 'scan_bwd', round(r['scan_bwd']['ms_per_step'],3), r['scan_bwd']['frac'], d['clocks']['sm_mhz'])"; }
echo cb256; run
echo cb128; SC_B200_LIB=$PWD/statecatcher_b200/csrc/libstatecatcher_b200_cb128.so run
echo cb256; run
echo cb128; SC_B200_LIB=$PWD/statecatcher_b200/csrc/libstatecatcher_b200_cb128.so run
