#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_drivers_gpu.py -m gpu -q -x 2>&1 | tail -3 > $O/r2c48_tests.txt
cat $O/r2c48_tests.txt
timeout 600 python profiles/scripts/pack_bench.py > $O/r2c48_pack_bench.txt 2>&1
grep '"col_major": false' $O/r2c48_pack_bench.txt | cut -c1-220
