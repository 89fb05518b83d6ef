#!/bin/bash
O=gpurun_out
for sl in 64 512 2048 4096; do
QGTC_TC05_EPI_SLEEP=$sl timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active --clock-control none -k regex:gemm_tc05 -s 12 -c 4 --csv --log-file $O/r2c34_sleep_$sl.csv \
    python bench.py --steps 1 --warmup 3 --passes 1 --no-cpu-baseline --no-epoch --no-zero-tile > /dev/null 2>&1
echo "sleep $sl"; grep -v "^==" $O/r2c34_sleep_$sl.csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for r in rows[1:]: print(r[h.index('ID')], r[h.index('Kernel Name')][22:52], r[h.index('Metric Name')][:40], r[h.index('Metric Value')])
"
done
