#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_tc05_gpu.py tests/test_ops_gpu.py -m gpu -q -x 2>&1 | tail -3 > $O/r2c32_tests.txt
cat $O/r2c32_tests.txt
QGTC_TC05_HELP=0 timeout 300 python profiles/scripts/tc05_ab.py help0 2>&1 | grep '"bits": [48],' > $O/r2c32_ab.txt
timeout 300 python profiles/scripts/tc05_ab.py help1 2>&1 | grep '"bits": [48],' >> $O/r2c32_ab.txt
cat $O/r2c32_ab.txt
