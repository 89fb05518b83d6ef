#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "val2bit" 2>&1 | tail -3 > $O/r2c47_tests.txt
cat $O/r2c47_tests.txt
timeout 600 python profiles/scripts/pack_batched_bench.py > $O/r2c47_pack_batched.txt 2>&1
cat $O/r2c47_pack_batched.txt
