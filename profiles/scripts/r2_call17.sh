#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_tc05_gpu.py -m gpu -q -x 2>&1 | tail -5 > $O/r2c17_tests.txt
cat $O/r2c17_tests.txt
: > $O/r2c17_ab.txt
QGTC_TC05_FIXED=0 timeout 300 python profiles/scripts/tc05_ab.py fixed0 quick >> $O/r2c17_ab.txt 2>&1
timeout 300 python profiles/scripts/tc05_ab.py fixed1 quick >> $O/r2c17_ab.txt 2>&1
QGTC_TC05_EPI_SLEEP=2048 timeout 300 python profiles/scripts/tc05_ab.py fixed1_sleep2048 quick >> $O/r2c17_ab.txt 2>&1
QGTC_TC05_EPI_SLEEP=8192 timeout 300 python profiles/scripts/tc05_ab.py fixed1_sleep8192 quick >> $O/r2c17_ab.txt 2>&1
QGTC_TC05_EPI_SLEEP=64 timeout 300 python profiles/scripts/tc05_ab.py fixed1_sleep64 quick >> $O/r2c17_ab.txt 2>&1
cat $O/r2c17_ab.txt
