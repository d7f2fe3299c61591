mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "scan" > gpurun_out/c52_tests.log 2>&1; tail -n 3 gpurun_out/c52_tests.log
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline_by_kernel']
print('RESULT', round(d['ms_per_step'],2), 'scan_bwd', round(r['scan_bwd']['ms_per_step'],3), r['scan_bwd']['frac'], 'gemm', round(r['gemm']['ms_per_step'],2), d['clocks']['sm_mhz'])"; }
echo ck2; run
echo base; SC_B200_LIB=$PWD/statecatcher_b200/csrc/libsc_base.so run
echo ck2; run
echo base; SC_B200_LIB=$PWD/statecatcher_b200/csrc/libsc_base.so run
