#!/bin/bash
# Round-2 GPU call 2: the grouped engine -- parity first (bounded), then the whole suite, then first timings.
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_grp_gpu.py -q -x 2>&1 | tail -25 > $O/r2c2_grp_tests.txt
cat $O/r2c2_grp_tests.txt
python - <<'PY' > $O/r2c2_errflag.txt 2>&1
import ctypes
from qgtc_ppopp22_b200 import capi
L = capi.lib(); L.qgtc_debug_tc05_error_flag.restype = ctypes.c_uint
print("watchdog flag: 0x%x" % L.qgtc_debug_tc05_error_flag())
PY
if grep -q "failed\|error" $O/r2c2_grp_tests.txt; then echo "grouped tests failed: skipping the rest"; exit 1; fi
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/r2c2_pytest_gpu.txt
cat $O/r2c2_pytest_gpu.txt
timeout 600 python profiles/scripts/epoch_modes.py ppi 2 > $O/r2c2_epoch_ppi.txt 2>&1
timeout 900 python profiles/scripts/epoch_modes.py ogbn-arxiv 1 2 4 8 > $O/r2c2_epoch_arxiv.txt 2>&1
timeout 600 python profiles/scripts/epoch_modes.py gin 2 > $O/r2c2_epoch_gin.txt 2>&1
timeout 300 python profiles/scripts/single_call_sweep.py > $O/r2c2_single_call.txt 2>&1
cat $O/r2c2_epoch_*.txt $O/r2c2_single_call.txt
