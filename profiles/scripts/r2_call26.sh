#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_tc05_gpu.py -m gpu -q -x 2>&1 | tail -3 > $O/r2c26_tests.txt
cat $O/r2c26_tests.txt
timeout 300 python profiles/scripts/tc05_ab.py own_full > $O/r2c26_ab.txt 2>&1
cat $O/r2c26_ab.txt
QGTC_TC05_TRACE=1 QGTC_TC05_TRACE_F4=1 timeout 120 python profiles/scripts/tc05_trace.py 64 > $O/r2c26_trace_f4_n64.txt 2>&1
sed -n 1,32p $O/r2c26_trace_f4_n64.txt
