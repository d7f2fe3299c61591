mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ctc_head.py tests/test_gpu_kernels.py tests/test_gpu_configs1_parity.py tests/test_gpu_zz_ctc_forms.py tests/test_gpu_zglue_golden.py -x -q -k "ctc or head or compute_loss or repeated" > gpurun_out/c47_tests.log 2>&1; tail -n 5 gpurun_out/c47_tests.log
timeout 120 python profiles/ctc_time.py | tee gpurun_out/c47_ctc_time.txt
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline_by_kernel']
print('RESULT', round(d['ms_per_step'],2), 'ctc', r['ctc']['ms_per_step'], r['ctc']['frac'], {k: r[k]['ms_per_step'] for k in r if k.startswith('ctc_')}, d['clocks']['sm_mhz'])"; }
run; run
