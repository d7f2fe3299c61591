set -x
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 15 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('RESULT', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['clocks'])"; }
python bench.py --steps 15 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('RESULT1', d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['clocks'])"
SC_DP_OVERLAP=1 run
SC_DP_OVERLAP=0 run
SC_DP_OVERLAP=1 NCCL_MAX_CTAS=4 run
SC_DP_OVERLAP=0 NCCL_MAX_CTAS=32 NCCL_MIN_CTAS=32 run
SC_DP_OVERLAP=1 run
