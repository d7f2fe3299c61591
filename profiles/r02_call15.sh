mkdir -p gpurun_out
timeout 120 python profiles/ctc_ws_debug.py > gpurun_out/c15_dbg.txt 2>&1
timeout 120 python profiles/ctc_time.py > gpurun_out/c15_ctc_time.txt 2>&1
cat gpurun_out/c15_dbg.txt gpurun_out/c15_ctc_time.txt
