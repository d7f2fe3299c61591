mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "layernorm" 2>&1 | tail -15
SC_LN_BWD_STAGE=0 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "layernorm" 2>&1 | tail -3
