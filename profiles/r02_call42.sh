mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ctc_head.py tests/test_gpu_kernels.py tests/test_gpu_configs1_parity.py tests/test_gpu_zz_ctc_forms.py tests/test_gpu_zglue_golden.py tests/test_gpu_shim_route.py -x -q -k "ctc or head or compute_loss or shim or scaler or register or repeated" > gpurun_out/c42_tests.log 2>&1; tail -n 6 gpurun_out/c42_tests.log
timeout 120 python profiles/ctc_time.py | tee gpurun_out/c42_ctc_time.txt
timeout 300 python profiles/ctc_head_exp.py > gpurun_out/c42_head_exp.txt 2>&1; cat gpurun_out/c42_head_exp.txt
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline_by_kernel']
print('RESULT', round(d['ms_per_step'],2), 'ctc', r['ctc']['ms_per_step'], r['ctc']['frac'], {k: r[k]['ms_per_step'] for k in r if k.startswith('ctc_')}, d['clocks']['sm_mhz'])"; }
run; run
echo overlap4; SC_CTC_OVERLAP=1 SC_CTC_PHASES=4 run
