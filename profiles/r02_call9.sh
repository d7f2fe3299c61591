mkdir -p gpurun_out
python profiles/ctc_only.py 2 > gpurun_out/c12_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"ctc_lin64_kernel" -c 1 -o gpurun_out/c12_prof_lin64 python profiles/ctc_only.py 2 > gpurun_out/c12_ncu.log 2>&1
tail -2 gpurun_out/c12_ncu.log
