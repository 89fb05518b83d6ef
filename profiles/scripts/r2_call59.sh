#!/bin/bash
O=gpurun_out
for gsz in 0 16 8 4 2; do
if [ $gsz = 0 ]; then unset QGTC_HOST_GROUP; else export QGTC_HOST_GROUP=$gsz; fi
timeout 300 python bench.py --steps 3 --warmup 3 --no-zero-tile --no-cpu-baseline --no-epoch > $O/r2c59_bench_g$gsz.json 2> $O/r2c59_bench_g$gsz.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c59_bench_g$gsz.json").read().strip().splitlines()[-1])
print("host group=$gsz e2e", d["e2e"]["value"], d["e2e"]["matches_device_result"])
PY
done
