mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ctc_head.py tests/test_gpu_kernels.py tests/test_gpu_configs1_parity.py tests/test_gpu_zz_ctc_forms.py tests/test_gpu_zglue_golden.py -x -q -k "ctc or head or compute_loss or repeated" > gpurun_out/c48_tests.log 2>&1; tail -n 5 gpurun_out/c48_tests.log
timeout 120 python profiles/ctc_time.py | tee gpurun_out/c48_ctc_time.txt
