#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_tc05_gpu.py tests/test_drivers_gpu.py tests/test_ops_gpu.py -m gpu -q -x 2>&1 | tail -3 > $O/r2c43_tests.txt
cat $O/r2c43_tests.txt
QGTC_TC05_PREP_PDL=0 timeout 300 python profiles/scripts/single_call_sweep.py 2>&1 | sed 's/^/pdl0 /' > $O/r2c43_single_call.txt
timeout 300 python profiles/scripts/single_call_sweep.py 2>&1 | sed 's/^/pdl1 /' >> $O/r2c43_single_call.txt
cat $O/r2c43_single_call.txt
for pdl in 0 1; do
QGTC_TC05_PREP_PDL=$pdl timeout 600 python bench.py --steps 10 --warmup 5 --no-zero-tile --no-cpu-baseline --no-epoch > $O/r2c43_bench_pdl$pdl.json 2> $O/r2c43_bench_pdl$pdl.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c43_bench_pdl$pdl.json").read().strip().splitlines()[-1])
print("pdl=$pdl value", d["value"], d["per_bit_tops_per_call"], d["single_gemm_per_launch_tops"], d["clocks"]["sm_mhz"], d["parity_spot_check"])
PY
done
