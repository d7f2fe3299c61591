run() { python bench.py --steps 15 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline_by_kernel']
print('RESULT', round(d['ms_per_step'],2), 'gemm', round(r['gemm']['ms_per_step'],3), 'frac', round(r['gemm']['frac'],3), d['clocks']['sm_mhz'])"; }
echo default; run
echo multicast; SC_GEMM_PAIR=0 run
echo pair_all; SC_GEMM_PAIR=1 run
echo default; run
