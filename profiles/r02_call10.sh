mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "ctc" > gpurun_out/c14_ctc_tests.log 2>&1
timeout 600 python -m pytest tests/test_gpu_configs1_parity.py -q -k "ctc" > gpurun_out/c14_parity.log 2>&1
timeout 120 python profiles/ctc_time.py > gpurun_out/c14_ctc_time_lin.txt 2>&1
python profiles/ctc_only.py 3 > gpurun_out/c14_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__cycles_elapsed.max --clock-control none -k regex:"ctc_" -c 12 --csv --log-file gpurun_out/c14_launches.csv python profiles/ctc_only.py 3 > gpurun_out/c14_ncu.log 2>&1
tail -n 4 gpurun_out/c14_ctc_tests.log gpurun_out/c14_parity.log; cat gpurun_out/c14_ctc_time_*.txt
grep -E "gpu__time_duration" gpurun_out/c14_launches.csv | awk -F'","' '{print substr($5,1,40), $NF}' | tail -8
