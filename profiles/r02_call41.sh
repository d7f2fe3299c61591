mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ctc_head.py tests/test_gpu_kernels.py tests/test_gpu_configs1_parity.py tests/test_gpu_zz_ctc_forms.py tests/test_gpu_zglue_golden.py tests/test_gpu_shim_route.py tests/test_gpu_module.py tests/test_gpu_pipeline.py tests/test_gpu_zzz_graph_train.py -x -q > gpurun_out/c41_tests.log 2>&1; tail -n 12 gpurun_out/c41_tests.log
for u in 31 95 150 255; do echo "umax $u"; CTC_UMAX=$u timeout 120 python profiles/ctc_time.py 2>&1 | grep -v loss; done | tee gpurun_out/c41_ctc_time_widths.txt
timeout 120 python profiles/ctc_time.py | tee gpurun_out/c41_ctc_time.txt
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline_by_kernel']
print('RESULT', round(d['ms_per_step'],2), 'ctc', r['ctc']['ms_per_step'], r['ctc']['frac'], {k: r[k]['ms_per_step'] for k in r if k.startswith('ctc_')}, d['clocks']['sm_mhz'])"; }
run; run
