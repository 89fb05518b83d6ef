#!/bin/bash
# call 63: driver tests after ClusterIter's dense route moved to the batched pack
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_drivers_gpu.py tests/test_reference_gpu.py -m gpu -x -q > gpurun_out/r2c63_pytest.txt 2>&1
tail -3 gpurun_out/r2c63_pytest.txt
