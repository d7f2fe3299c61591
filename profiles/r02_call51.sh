mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_configs1_parity.py tests/test_gpu_module.py -q -k "scan or module or golden or configs" > gpurun_out/c51_tests.log 2>&1; tail -n 5 gpurun_out/c51_tests.log
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline_by_kernel']
print('RESULT', round(d['ms_per_step'],2), 'scan_fwd', round(r['scan_fwd']['ms_per_step'],3), r['scan_fwd']['frac'], 'scan_bwd', round(r['scan_bwd']['ms_per_step'],3), r['scan_bwd']['frac'], 'gemm', round(r['gemm']['ms_per_step'],2), d['clocks']['sm_mhz'])"; }
echo ws; run
echo barrier; SC_SCAN_BWD_WS=0 run
echo ws; run
echo barrier; SC_SCAN_BWD_WS=0 run
