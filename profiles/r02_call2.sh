mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "ctc" > gpurun_out/c2_ctc_tests.log 2>&1
timeout 600 python -m pytest tests/test_gpu_configs1_parity.py tests/test_gpu_shim_route.py tests/test_gpu_zz_ctc_forms.py tests/test_gpu_zglue_golden.py -q > gpurun_out/c2_parity.log 2>&1
timeout 120 python profiles/ctc_time.py > gpurun_out/c2_ctc_time_lin.txt 2>&1
SC_CTC_LIN=0 timeout 120 python profiles/ctc_time.py > gpurun_out/c2_ctc_time_log.txt 2>&1
SC_CTC_FORCE_LOSSY=1 timeout 120 python profiles/ctc_time.py > gpurun_out/c2_ctc_time_forced_repair.txt 2>&1
timeout 300 python bench.py --workload cfg1 --steps 10 --warmup 3 --detail --no-cpu-baseline > gpurun_out/c2_bench_cfg1.json 2> gpurun_out/c2_bench_cfg1_detail.txt
tail -5 gpurun_out/c2_ctc_tests.log gpurun_out/c2_parity.log; cat gpurun_out/c2_ctc_time_*.txt
