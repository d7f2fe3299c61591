mkdir -p gpurun_out
for v in BASE ONE_STORE ONE_LOAD BOTH; do for u in 31 150; do echo "variant $v umax $u"; CTC_UMAX=$u SC_B200_LIB=$PWD/statecatcher_b200/csrc/libsc_dbg_$v.so timeout 120 python profiles/ctc_time.py 2>&1 | grep lattice; done; done | tee gpurun_out/c40_lattice_io_experiment.txt
