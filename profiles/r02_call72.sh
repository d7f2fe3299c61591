mkdir -p gpurun_out
# final-code evidence: GPU suite, smoke, default bench + reference arm, launch list of the bench command, --set full of the first GEMMs
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c72_gpu_suite.log 2>&1; tail -n 6 gpurun_out/c72_gpu_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/c72_smoke.log 2>&1; tail -n 2 gpurun_out/c72_smoke.log
timeout 600 python bench.py > gpurun_out/c72_bench_default.json 2> gpurun_out/c72_bench_default.err; tail -c 1600 gpurun_out/c72_bench_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c72_bench_reference.json 2> gpurun_out/c72_bench_reference.err; tail -c 300 gpurun_out/c72_bench_reference.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c72_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/c72_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c72_ncu_list.log 2>&1
python profiles/ncu_launch_shares.py gpurun_out/c72_launches.csv > gpurun_out/c72_launch_shares.txt 2>&1; head -12 gpurun_out/c72_launch_shares.txt
python profiles/run_step.py --steps 1 > gpurun_out/c72_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc" -c 6 -o gpurun_out/c72_prof_gemm -f python profiles/run_step.py --steps 1 > gpurun_out/c72_ncu_full.log 2>&1
ncu -i gpurun_out/c72_prof_gemm.ncu-rep --page raw --csv > gpurun_out/c72_gemm_raw.csv 2>/dev/null; python profiles/ncu_summary.py gpurun_out/c72_gemm_raw.csv > gpurun_out/c72_gemm_summary.txt; grep -E "Kernel Name|time_duration|dram__bytes|tensor_cycles" gpurun_out/c72_gemm_summary.txt | head -40
rm -f gpurun_out/c72_prof_gemm.ncu-rep
