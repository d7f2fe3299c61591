mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c20_gpu_suite.log 2>&1
tail -n 8 gpurun_out/c20_gpu_suite.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/parity_configs1.json'))['module_f32']
print({k:v for k,v in d.items() if not k.startswith('grad_rel_l2/')})
PY
