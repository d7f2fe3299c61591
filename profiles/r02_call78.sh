mkdir -p gpurun_out
# bf16 column sums with 16-byte loads (output_proj bias gradient): tests, then the cfg2 step's colsum line
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_module.py tests/test_gpu_rnnt.py -q -k "colsum or golden or module or joint_bwd or fused_head" 2>&1 | tail -3
timeout 600 python bench.py --detail --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c78.json 2> gpurun_out/c78_detail.txt
grep -E "sc_colsum" gpurun_out/c78_detail.txt | tail -2
python -c "
import json; d=json.loads(open('gpurun_out/c78.json').read().strip().split(chr(10))[-1]); print('ms/step', round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
