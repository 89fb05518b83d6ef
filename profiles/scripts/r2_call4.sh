#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_grp_gpu.py -q -x 2>&1 | tail -5 > $O/r2c4_grp_tests.txt
cat $O/r2c4_grp_tests.txt
if grep -q "failed\|error" $O/r2c4_grp_tests.txt; then exit 1; fi
timeout 120 python profiles/scripts/grp_trace.py agg > $O/r2c4_trace_agg.txt 2>&1
timeout 120 python profiles/scripts/grp_trace.py xw > $O/r2c4_trace_xw.txt 2>&1
timeout 600 python profiles/scripts/epoch_modes.py ppi 2 > $O/r2c4_epoch_ppi.txt 2>&1
cat $O/r2c4_epoch_ppi.txt; cat $O/r2c4_trace_agg.txt
