run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline_by_kernel']
print('RESULT', round(d['ms_per_step'],2), 'scan_fwd', round(r['scan_fwd']['ms_per_step'],3), r['scan_fwd']['frac'], 'scan_bwd', round(r['scan_bwd']['ms_per_step'],3), r['scan_bwd']['frac'], 'gemm', round(r['gemm']['ms_per_step'],2), d['clocks']['sm_mhz'])"; }
echo onebox; run
echo fiveboxes; SC_B200_LIB=$PWD/statecatcher_b200/csrc/libsc_old_scan.so run
echo onebox; run
echo fiveboxes; SC_B200_LIB=$PWD/statecatcher_b200/csrc/libsc_old_scan.so run
