python tools/c5_step.py > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 620 -c 24 --csv --log-file gpurun_out/r2_c5_launches.csv python tools/c5_step.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_c5_launches.csv')) if len(r)>5]
h=rows[0]; ik=h.index('Kernel Name'); iv=h.index('Metric Value')
for r in rows[1:]: print(r[ik][:60], r[iv])
PY
NCU="ncu --set full --clock-control none --import-source on"
python tools/kbench2.py c3 > gpurun_out/plain_c3.log 2>&1 && $NCU -k regex:worldline_smem_table -s 20 -c 1 -o gpurun_out/r2_c3 python tools/kbench2.py c3 > gpurun_out/ncu_c3.log 2>&1
python tools/ncu_summary.py gpurun_out/r2_c3.ncu-rep 4194304 > gpurun_out/r2_c3_summary.txt 2>&1
NCU_LINES_TOP=60 python tools/ncu_lines.py gpurun_out/r2_c3.ncu-rep worldline_smem_table_kernelILi0ELi64 4194304 > gpurun_out/r2_c3_lines.txt 2>&1
ncu -i gpurun_out/r2_c3.ncu-rep --page raw --csv > gpurun_out/r2_c3_raw.csv 2>/dev/null
rm -f gpurun_out/r2_c3.ncu-rep
head -12 gpurun_out/plain_c3.log
