mkdir -p gpurun_out
for d in 0 1 2 4 8 3 7 15; do echo "dbg $d"; SC_CTC_DBG=$d timeout 120 python profiles/ctc_time.py 2>&1 | grep lattice; done > gpurun_out/c6_dbg.txt 2>&1
cat gpurun_out/c6_dbg.txt
