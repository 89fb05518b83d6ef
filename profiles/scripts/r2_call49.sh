#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_tc05_gpu.py -m gpu -q -x 2>&1 | tail -3 > $O/r2c49_tests.txt
cat $O/r2c49_tests.txt
QGTC_TC05_FIXED=1 timeout 300 python profiles/scripts/tc05_ab.py fixed_n128 2>&1 | grep '"N": 128' > $O/r2c49_ab.txt
cat $O/r2c49_ab.txt
