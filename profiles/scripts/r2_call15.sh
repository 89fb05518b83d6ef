#!/bin/bash
O=gpurun_out
mkdir -p $O
QGTC_TC05_SS=1 QGTC_TC05_FP4=0 timeout 600 python -m pytest tests/test_tc05_gpu.py -m gpu -q -x -k "bitmm2bit_and_int or headline or batched or ragged or fig8a or tilemap" 2>&1 | tail -5 > $O/r2c15_ss_tests.txt
cat $O/r2c15_ss_tests.txt
python - <<'PY'
import ctypes
from qgtc_ppopp22_b200 import capi
L = capi.lib(); L.qgtc_debug_tc05_error_flag.restype = ctypes.c_uint
print("watchdog flag: 0x%x" % L.qgtc_debug_tc05_error_flag())
PY
if grep -q "failed\|error" $O/r2c15_ss_tests.txt; then exit 1; fi
QGTC_TC05_SS=0 QGTC_TC05_FP4=0 timeout 300 python profiles/scripts/ss_sweep.py > $O/r2c15_ss0.txt 2>&1
QGTC_TC05_SS=1 QGTC_TC05_FP4=0 timeout 300 python profiles/scripts/ss_sweep.py > $O/r2c15_ss1.txt 2>&1
paste -d'\n' $O/r2c15_ss0.txt $O/r2c15_ss1.txt | cut -c1-200
