mkdir -p gpurun_out
SC_B200_LIB=$PWD/statecatcher_b200/csrc/libsc_eb3.so timeout 600 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_kernels.py -q -k "gemm" > gpurun_out/c54_tests.log 2>&1; tail -n 3 gpurun_out/c54_tests.log
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline_by_kernel']
print('RESULT', round(d['ms_per_step'],2), 'gemm', round(r['gemm']['ms_per_step'],3), r['gemm']['frac'], d['clocks']['sm_mhz'])"; }
echo eb2; run
echo eb3; SC_B200_LIB=$PWD/statecatcher_b200/csrc/libsc_eb3.so run
echo eb2; run
echo eb3; SC_B200_LIB=$PWD/statecatcher_b200/csrc/libsc_eb3.so run
