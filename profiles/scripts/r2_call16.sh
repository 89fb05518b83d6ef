#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_reference_gpu.py tests/test_drivers_gpu.py -m gpu -q -x 2>&1 | tail -5 > $O/r2c16_tests.txt
cat $O/r2c16_tests.txt
timeout 120 python profiles/scripts/grp_trace.py agg 759 20 2 16 > $O/r2c16_trace_agg_1unit.txt 2>&1
cat $O/r2c16_trace_agg_1unit.txt | head -60
