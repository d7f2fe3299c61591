mkdir -p gpurun_out
./profiles/micro/tput_bench > gpurun_out/c3_tput_bench.txt 2>&1
python profiles/ctc_only.py 2 > gpurun_out/c3_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"ctc_lin64|ctc_grad|ctc_lse" -c 8 -o gpurun_out/c3_prof_ctc python profiles/ctc_only.py 2 > gpurun_out/c3_ncu.log 2>&1
cat gpurun_out/c3_tput_bench.txt; tail -3 gpurun_out/c3_ncu.log
