#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_grp_gpu.py -q -x 2>&1 | tail -5 > $O/r2c6_grp_tests.txt
cat $O/r2c6_grp_tests.txt
if grep -q "failed\|error" $O/r2c6_grp_tests.txt; then exit 1; fi
timeout 120 python profiles/scripts/grp_trace.py agg > $O/r2c6_trace_agg.txt 2>&1
timeout 120 python profiles/scripts/grp_trace.py xw > $O/r2c6_trace_xw.txt 2>&1
timeout 600 python profiles/scripts/epoch_modes.py ppi 2 > $O/r2c6_epoch_ppi.txt 2>&1
timeout 600 python profiles/scripts/epoch_modes.py ogbn-arxiv 2 8 > $O/r2c6_epoch_arxiv.txt 2>&1
timeout 600 python profiles/scripts/epoch_modes.py gin 2 > $O/r2c6_epoch_gin.txt 2>&1
timeout 300 python profiles/scripts/single_call_sweep.py > $O/r2c6_single_call.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > $O/r2c6_pytest_gpu.txt
cat $O/r2c6_epoch_*.txt $O/r2c6_single_call.txt $O/r2c6_pytest_gpu.txt; head -60 $O/r2c6_trace_xw.txt
