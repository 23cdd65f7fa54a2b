#!/bin/bash
mkdir -p gpurun_out
timeout -k 5 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'native|spmv_sjds|sjds_block' -s 4 -c 4 --csv --log-file gpurun_out/r02g_mv_launches.csv \
   python scripts/mv_ncu_target.py hubbard4x4 3 > gpurun_out/r02g_mv.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02g_mv_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows: print(r[0], r[4][:60], r[-3], r[-2], r[-1])
PY
