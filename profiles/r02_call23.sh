mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ctc_head.py -x -q > gpurun_out/c23_head_tests.log 2>&1; tail -n 25 gpurun_out/c23_head_tests.log
for p in 0 2 4 6 8 12; do SC_CTC_PHASES=$p timeout 120 python profiles/ctc_time.py > gpurun_out/c23_ctc_time_p$p.txt 2>&1; echo "phases $p"; cat gpurun_out/c23_ctc_time_p$p.txt; done
SC_CTC_OVERLAP=0 timeout 120 python profiles/ctc_time.py > gpurun_out/c23_ctc_time_seq.txt 2>&1; cat gpurun_out/c23_ctc_time_seq.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_configs1_parity.py tests/test_gpu_zz_ctc_forms.py tests/test_gpu_zglue_golden.py tests/test_gpu_shim_route.py -q -k "ctc or compute_loss or shim or scaler" > gpurun_out/c23_ctc_tests.log 2>&1; tail -n 8 gpurun_out/c23_ctc_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c23_n1.json 2> gpurun_out/c23_n1.err; tail -c 1800 gpurun_out/c23_n1.json
