mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c60_gpu_suite.log 2>&1; tail -n 4 gpurun_out/c60_gpu_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c60_smoke.log 2>&1; tail -n 1 gpurun_out/c60_smoke.log
timeout 600 python bench.py > gpurun_out/c60_bench_default.json 2> gpurun_out/c60_bench_default.err; tail -c 1500 gpurun_out/c60_bench_default.json
timeout 600 python bench.py --layer-norm --detail --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c60_ln.json 2> gpurun_out/c60_ln_detail.txt; python -c "
import json; d=json.loads(open('gpurun_out/c60_ln.json').read().strip().split(chr(10))[-1]); print('LN ms/step', d['ms_per_step'], d['clocks'])"
