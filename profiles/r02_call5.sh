mkdir -p gpurun_out
python profiles/ctc_only.py 3 > gpurun_out/c5_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__cycles_elapsed.max --clock-control none -k regex:"ctc_" --csv --log-file gpurun_out/c5_launches.csv python profiles/ctc_only.py 3 > gpurun_out/c5_ncu.log 2>&1
grep -v "^==" gpurun_out/c5_launches.csv | cut -d, -f5,13- | tail -40
