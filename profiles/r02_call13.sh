mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/c13_gpu_suite.log 2>&1
timeout 120 python profiles/ctc_time.py > gpurun_out/c13_ctc_time.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err
tail -n 6 gpurun_out/c13_gpu_suite.log; cat gpurun_out/c13_ctc_time.txt; tail -c 1800 gpurun_out/c13_bench.json
