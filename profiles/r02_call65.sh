mkdir -p gpurun_out
timeout 600 python bench.py --workload cfg5_fwd_b64 --detail --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c65_fwd.json 2> gpurun_out/c65_fwd_detail.txt
grep "sc_gemm_fwd\|sc_lucy_scan_fwd" gpurun_out/c65_fwd_detail.txt | head -16
python profiles/gemm_exp.py 2>&1 | tail -5
