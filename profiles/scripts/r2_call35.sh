#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_tc05_gpu.py -m gpu -q -x 2>&1 | tail -3 > $O/r2c35_tests.txt
cat $O/r2c35_tests.txt
timeout 300 python profiles/scripts/tc05_ab.py epi_poll1 > $O/r2c35_ab.txt 2>&1
cat $O/r2c35_ab.txt
