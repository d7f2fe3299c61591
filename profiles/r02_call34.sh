mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_module.py tests/test_gpu_kernels.py tests/test_gpu_zglue_golden.py tests/test_gpu_configs1_parity.py -q > gpurun_out/c34_tests.log 2>&1; tail -n 15 gpurun_out/c34_tests.log
