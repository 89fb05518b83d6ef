#!/bin/bash
# tcgen05 engine with 1..4 MMA issuer warps (QGTC_TC05_ISSUERS): batched 4096x4096xN, 2-bit features
for ni in 1 2; do
  echo "issuers=$ni"
  QGTC_TC05_ISSUERS=$ni python - <<'PY'
import sys, os, json
sys.path.insert(0, os.path.join(os.getcwd(), "profiles", "scripts"))
from engine_sweep import run
from qgtc_ppopp22_b200 import capi
for N in (64, 16, 128):
    ms, tops = run(capi.ENGINE_TCGEN05, 4096, 4096, N, 2, 128)
    print(json.dumps({"N": N, "ms": round(ms, 4), "tops": round(tops, 1)}))
PY
done
