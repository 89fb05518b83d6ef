#!/bin/bash
# 4-bit operand path: chunks per pipeline item (QGTC_TC05_CPS4), batched 4096x4096xN, 2-bit features
for c in 7 5 4 3; do echo "cps4=$c"; QGTC_TC05_CPS4=$c python - <<'PY'
import sys, os, json
sys.path.insert(0, os.path.join(os.getcwd(), "profiles", "scripts"))
from engine_sweep import run
from qgtc_ppopp22_b200 import capi
for N in (64, 16, 128):
    ms, tops = run(capi.ENGINE_TCGEN05, 4096, 4096, N, 2, 128)
    print(json.dumps({"N": N, "ms": round(ms, 4), "tops": round(tops, 1)}))
PY
done
