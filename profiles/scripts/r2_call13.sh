#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_grp_gpu.py tests/test_drivers_gpu.py -m gpu -q -x 2>&1 | tail -5 > $O/r2c13_tests.txt
cat $O/r2c13_tests.txt
if grep -q "failed\|error" $O/r2c13_tests.txt; then exit 1; fi
timeout 600 python profiles/scripts/epoch_modes.py ppi 2 > $O/r2c13_epoch_ppi.txt 2>&1
timeout 600 python profiles/scripts/epoch_modes.py ogbn-arxiv 2 8 > $O/r2c13_epoch_arxiv.txt 2>&1
timeout 600 python profiles/scripts/epoch_modes.py gin 2 > $O/r2c13_epoch_gin.txt 2>&1
cat $O/r2c13_epoch_*.txt
timeout 120 python profiles/scripts/grp_trace.py agg > $O/r2c13_trace_agg.txt 2>&1
head -64 $O/r2c13_trace_agg.txt | tail -52
