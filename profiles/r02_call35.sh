mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ctc_head.py tests/test_gpu_kernels.py tests/test_gpu_configs1_parity.py tests/test_gpu_zz_ctc_forms.py tests/test_gpu_zglue_golden.py tests/test_gpu_shim_route.py -q -k "ctc or compute_loss or shim or scaler or head or register" > gpurun_out/c35_ctc_tests.log 2>&1; tail -n 12 gpurun_out/c35_ctc_tests.log
timeout 120 python profiles/ctc_time.py > gpurun_out/c35_ctc_time.txt 2>&1; cat gpurun_out/c35_ctc_time.txt
SC_CTC_GRAD_REG=0 timeout 120 python profiles/ctc_time.py > gpurun_out/c35_ctc_time_smemrow.txt 2>&1; cat gpurun_out/c35_ctc_time_smemrow.txt
timeout 300 python profiles/ctc_head_exp.py > gpurun_out/c35_head_exp.txt 2>&1; cat gpurun_out/c35_head_exp.txt
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline_by_kernel']
print('RESULT', round(d['ms_per_step'],2), 'ctc', r['ctc']['ms_per_step'], r['ctc']['frac'], {k: r[k]['ms_per_step'] for k in r if k.startswith('ctc_')}, d['clocks']['sm_mhz'])"; }
echo reg; run
echo smemrow; SC_CTC_GRAD_REG=0 run
echo reg+overlap4; SC_CTC_OVERLAP=1 SC_CTC_PHASES=4 run
echo reg+overlap6; SC_CTC_OVERLAP=1 SC_CTC_PHASES=6 run
