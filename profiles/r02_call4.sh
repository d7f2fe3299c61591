mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "ctc" > gpurun_out/c4_ctc_tests.log 2>&1
timeout 600 python -m pytest tests/test_gpu_configs1_parity.py -q -k "ctc" > gpurun_out/c4_parity.log 2>&1
timeout 120 python profiles/ctc_time.py > gpurun_out/c4_ctc_time_lin.txt 2>&1
tail -n 5 gpurun_out/c4_ctc_tests.log gpurun_out/c4_parity.log; cat gpurun_out/c4_ctc_time_*.txt
