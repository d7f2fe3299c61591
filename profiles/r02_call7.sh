mkdir -p gpurun_out
python profiles/ctc_only.py 2 > gpurun_out/c7_plain.log 2>&1 || exit 1
for d in 0 1 2 4 8 15; do
  SC_CTC_DBG=$d ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"ctc_lin64_kernel" -c 2 --csv --log-file gpurun_out/c7_dbg$d.csv python profiles/ctc_only.py 2 > /dev/null 2>&1
  echo "dbg $d: $(grep -E 'gpu__time_duration|inst_executed' gpurun_out/c7_dbg$d.csv | tail -2 | awk -F'\",\"' '{print $(NF-2), $NF}' | tr '\n' ' ')"
done > gpurun_out/c7_dbg.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:"ctc_lin64_kernel" -c 1 -o gpurun_out/c7_prof_lin64 python profiles/ctc_only.py 2 > gpurun_out/c7_ncu.log 2>&1
cat gpurun_out/c7_dbg.txt
