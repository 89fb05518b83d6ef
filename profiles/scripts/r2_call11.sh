#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_grp_gpu.py -q -x -k "fused" 2>&1 | tail -8 > $O/r2c11_fused_tests.txt
cat $O/r2c11_fused_tests.txt
python - <<'PY'
import ctypes
from qgtc_ppopp22_b200 import capi
L = capi.lib(); L.qgtc_debug_tc05_error_flag.restype = ctypes.c_uint
print("watchdog flag: 0x%x" % L.qgtc_debug_tc05_error_flag())
PY
if grep -q "failed\|error" $O/r2c11_fused_tests.txt; then exit 1; fi
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $O/r2c11_pytest_gpu.txt
cat $O/r2c11_pytest_gpu.txt
if grep -q "failed\|error" $O/r2c11_pytest_gpu.txt; then exit 1; fi
timeout 600 python profiles/scripts/epoch_modes.py ppi 2 > $O/r2c11_epoch_ppi.txt 2>&1
timeout 600 python profiles/scripts/epoch_modes.py ogbn-arxiv 2 8 > $O/r2c11_epoch_arxiv.txt 2>&1
timeout 600 python profiles/scripts/epoch_modes.py gin 2 > $O/r2c11_epoch_gin.txt 2>&1
cat $O/r2c11_epoch_*.txt
