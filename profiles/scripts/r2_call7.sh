#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_grp_gpu.py -q -x 2>&1 | tail -5 > $O/r2c7_grp_tests.txt
cat $O/r2c7_grp_tests.txt
if grep -q "failed\|error" $O/r2c7_grp_tests.txt; then exit 1; fi
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > $O/r2c7_pytest_gpu.txt
timeout 600 python profiles/scripts/epoch_modes.py ppi 2 > $O/r2c7_epoch_ppi.txt 2>&1
timeout 600 python profiles/scripts/epoch_modes.py ogbn-arxiv 2 8 > $O/r2c7_epoch_arxiv.txt 2>&1
timeout 600 python profiles/scripts/epoch_modes.py gin 2 > $O/r2c7_epoch_gin.txt 2>&1
timeout 120 python profiles/scripts/grp_trace.py agg > $O/r2c7_trace_agg.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2c7_bench.json 2> $O/r2c7_bench.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > $O/r2c7_bench_ref.json 2> $O/r2c7_bench_ref.err
cat $O/r2c7_pytest_gpu.txt $O/r2c7_epoch_*.txt; tail -c 3000 $O/r2c7_bench.json; tail -5 $O/r2c7_bench.err; tail -c 1500 $O/r2c7_bench_ref.json; tail -5 $O/r2c7_bench_ref.err
