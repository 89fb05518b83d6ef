#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_tc05_gpu.py -m gpu -q -x 2>&1 | tail -3 > $O/r2c29_tests.txt
cat $O/r2c29_tests.txt
timeout 300 python profiles/scripts/tc05_ab.py split_u8 quick > $O/r2c29_ab.txt 2>&1
cat $O/r2c29_ab.txt
