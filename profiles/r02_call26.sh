mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ctc_head.py -x -q > gpurun_out/c26_head_tests.log 2>&1; tail -n 5 gpurun_out/c26_head_tests.log
timeout 300 python profiles/ctc_interference_exp.py > gpurun_out/c26_interference.txt 2>&1; cat gpurun_out/c26_interference.txt
timeout 300 python profiles/ctc_head_exp.py > gpurun_out/c26_head_exp.txt 2>&1; cat gpurun_out/c26_head_exp.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_configs1_parity.py tests/test_gpu_zz_ctc_forms.py tests/test_gpu_zglue_golden.py tests/test_gpu_shim_route.py -q -k "ctc or compute_loss or shim or scaler" > gpurun_out/c26_ctc_tests.log 2>&1; tail -n 8 gpurun_out/c26_ctc_tests.log
