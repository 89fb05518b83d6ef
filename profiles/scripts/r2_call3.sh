#!/bin/bash
# Round-2 GPU call 3: where does a grouped launch spend its time?  launch lists + one full capture per op kind
O=gpurun_out
mkdir -p $O
for cfg in "ppi 2" "ogbn-arxiv 8" "gin 2"; do
  set -- $cfg
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gemm_grp --csv --log-file $O/r2c3_launches_$1_$2.csv \
      python profiles/scripts/grouped_epoch_once.py $1 $2 2 > $O/r2c3_launches_$1_$2.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_grp -s 6 -c 6 -o $O/r2c3_prof_grp_ppi -f \
    python profiles/scripts/grouped_epoch_once.py ppi 2 2 > $O/r2c3_prof.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_grp -s 6 -c 6 -o $O/r2c3_prof_grp_arxiv8 -f \
    python profiles/scripts/grouped_epoch_once.py ogbn-arxiv 8 2 > $O/r2c3_prof2.log 2>&1
for f in $O/r2c3_launches_*.csv; do echo $f; tail -12 $f | cut -d, -f5,14- ; done
