#!/bin/bash
# Round-1 measurement batch on one B200 (run from the repo root); everything lands in gpurun_out/.
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/r1_pytest_gpu.txt
python bench.py --steps 20 --warmup 3 > $O/r1_bench.json 2> $O/r1_bench.err
python bench.py --impl reference --steps 5 --warmup 3 > $O/r1_bench_reference.json 2> $O/r1_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_|prep_b" -s 40 -c 80 --csv --log-file $O/r1_launches_bench.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-epoch > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc05 -s 14 -c 2 -o $O/r1_prof_bench_tc05 -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-epoch > /dev/null 2>&1
python profiles/scripts/engine_sweep.py > $O/r1_engine_sweep.txt 2>&1
python profiles/scripts/zero_tile_sweep.py > $O/r1_zero_tile_sweep.txt 2>&1
python profiles/scripts/epoch_sweep.py > $O/r1_epoch_sweep_n1.txt 2>&1
python profiles/scripts/single_call_sweep.py > $O/r1_single_call_sweep.txt 2>&1
bash profiles/scripts/amode_sweep.sh > $O/r1_amode_sweep.txt 2>&1
cat $O/r1_pytest_gpu.txt; tail -c 600 $O/r1_bench.json
