mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_configs1_parity.py tests/test_gpu_zz_ctc_forms.py tests/test_gpu_zglue_golden.py -q -k "ctc or compute_loss" > gpurun_out/c22_ctc_tests.log 2>&1
timeout 120 python profiles/ctc_time.py > gpurun_out/c22_ctc_time.txt 2>&1
SC_CTC_GRAD_PIPE=0 timeout 120 python profiles/ctc_time.py > gpurun_out/c22_ctc_time_nopipe.txt 2>&1
tail -n 4 gpurun_out/c22_ctc_tests.log; cat gpurun_out/c22_ctc_time.txt gpurun_out/c22_ctc_time_nopipe.txt
