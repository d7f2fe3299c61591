mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
./profiles/micro/lat_bench > gpurun_out/c1_lat_bench.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_configs1_parity.py -q > gpurun_out/c1_parity.log 2>&1
SC_RUN_EXPERIMENTAL=1 SC_OPT_VEC=1 timeout 600 python -m pytest tests/test_gpu_optim.py tests/test_gpu_zz_ctc_forms.py tests/test_gpu_zz_optim_ext.py tests/test_gpu_zzz_graph_train.py -q > gpurun_out/c1_new_tests.log 2>&1
SC_RUN_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_zzzz_ctc_linear.py -q > gpurun_out/c1_ctc_linear_test.log 2>&1
for w in 0 2 3 4; do SC_CTC_WAVE=$w timeout 120 python profiles/ctc_time.py > gpurun_out/c1_ctc_time_wave$w.txt 2>&1; done
timeout 200 python profiles/optim_time.py > gpurun_out/c1_optim_time_scalar.txt 2>&1
SC_OPT_VEC=1 timeout 200 python profiles/optim_time.py > gpurun_out/c1_optim_time_vec.txt 2>&1
timeout 200 python profiles/graph_train_time.py > gpurun_out/c1_graph_train_time.txt 2>&1
tail -3 gpurun_out/c1_*.log
