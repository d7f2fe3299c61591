mkdir -p gpurun_out
timeout 300 python profiles/ctc_interference_exp.py > gpurun_out/c25_interference.txt 2>&1; cat gpurun_out/c25_interference.txt
