mkdir -p gpurun_out
python profiles/ln_exp.py 2>&1 | tee gpurun_out/c70_ln_exp.txt
ncu --set full --clock-control none --import-source on -k regex:"layernorm" -s 8 -c 3 -o gpurun_out/c70_ln python profiles/ln_exp.py > gpurun_out/c70_ncu.log 2>&1
ncu -i gpurun_out/c70_ln.ncu-rep --page raw --csv > gpurun_out/c70_ln_raw.csv 2>/dev/null
python profiles/ncu_summary.py gpurun_out/c70_ln_raw.csv | tee gpurun_out/c70_ln_summary.txt | head -80
