run() { python bench.py --steps 15 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline_by_kernel']
print('RESULT', round(d['ms_per_step'],2), 'scan_fwd', round(r['scan_fwd']['ms_per_step'],3), 'scan_bwd', round(r['scan_bwd']['ms_per_step'],3), d['clocks']['sm_mhz'])"; }
echo default; run
echo VEC1; SC_SCAN_VEC=1 run
echo default; run
