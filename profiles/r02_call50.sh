mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c50_gpu_suite.log 2>&1; tail -n 6 gpurun_out/c50_gpu_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c50_smoke.log 2>&1; tail -n 2 gpurun_out/c50_smoke.log
timeout 600 python bench.py > gpurun_out/c50_bench_default.json 2> gpurun_out/c50_bench_default.err; tail -c 1600 gpurun_out/c50_bench_default.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c50_bench_reference.json 2> gpurun_out/c50_bench_reference.err; tail -c 400 gpurun_out/c50_bench_reference.json
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c50_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/c50_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/c50_ncu_list.log 2>&1
python profiles/run_step.py --steps 1 > gpurun_out/c50_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lucy_scan|ctc_" -c 20 -o gpurun_out/c50_prof_scan_ctc -f python profiles/run_step.py --steps 1 > gpurun_out/c50_ncu_full.log 2>&1
timeout 600 python bench.py --workload cfg1 --graph --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c50_cfg1_graph.json 2>/dev/null; tail -c 300 gpurun_out/c50_cfg1_graph.json
ls -la gpurun_out/c50_*
