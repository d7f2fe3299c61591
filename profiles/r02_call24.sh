mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ctc_head.py -x -q > gpurun_out/c24_head_tests.log 2>&1; tail -n 5 gpurun_out/c24_head_tests.log
timeout 300 python profiles/ctc_head_exp.py > gpurun_out/c24_head_exp.txt 2>&1; cat gpurun_out/c24_head_exp.txt
