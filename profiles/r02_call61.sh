mkdir -p gpurun_out
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline $1 2>gpurun_out/c61.err | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('RESULT $1', round(d['ms_per_step'],3), round(d['value']), 'e2e', round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'], d['config']['cuda_graph'])"; tail -2 gpurun_out/c61.err; }
run ""; run "--graph"; run ""; run "--graph"
