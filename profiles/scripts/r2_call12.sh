#!/bin/bash
# evidence run: launch list of a whole ppi epoch (every kernel), full ncu capture of the grouped kernels and of gemm_tc05 (traffic), final bench N=1 both arms
O=gpurun_out
mkdir -p $O
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2c12_launches_epoch_ppi.csv \
    python profiles/scripts/grouped_epoch_once.py ppi 2 2 > $O/r2c12_launches_epoch_ppi.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2c12_launches_epoch_gin.csv \
    python profiles/scripts/grouped_epoch_once.py gin 2 2 > $O/r2c12_launches_epoch_gin.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_grp -s 4 -c 4 -o $O/r2c12_prof_grp_ppi -f \
    python profiles/scripts/grouped_epoch_once.py ppi 2 2 > $O/r2c12_prof.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:gemm_tc05 -s 12 -c 4 -o $O/r2c12_prof_tc05_bench -f \
    python bench.py --steps 1 --warmup 3 --passes 1 --no-cpu-baseline --no-epoch --no-zero-tile > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_|prep_b" -s 24 -c 64 --csv --log-file $O/r2c12_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --passes 1 --no-cpu-baseline --no-epoch --no-zero-tile > /dev/null 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2c12_bench.json 2> $O/r2c12_bench.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > $O/r2c12_bench_ref.json 2> $O/r2c12_bench_ref.err
tail -c 1500 $O/r2c12_bench.json; tail -3 $O/r2c12_bench.err; tail -c 600 $O/r2c12_bench_ref.json
cut -d, -f5,14- $O/r2c12_launches_epoch_ppi.csv | tail -14
