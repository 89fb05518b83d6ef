#!/bin/bash
# call 73: grouped engine, converter half-groups on alternate items (<= 64 conversion tasks per item): parity, fuzz, epoch A/B
O=gpurun_out
timeout 900 python -m pytest tests/test_grp_gpu.py tests/test_drivers_gpu.py tests/test_reference_gpu.py -m gpu -q -x 2>&1 | tail -3 > $O/r2c73_tests.txt
cat $O/r2c73_tests.txt
timeout 200 python profiles/scripts/fuzz_grp.py 40 5 2>&1 | tail -3
for tv in 0 1 1; do
QGTC_GRP_CONV_ALT=$tv timeout 600 python bench.py --steps 3 --warmup 3 --no-zero-tile --no-cpu-baseline > $O/r2c73_bench_t$tv.json 2> $O/r2c73_bench_t$tv.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c73_bench_t$tv.json").read().strip().splitlines()[-1])
e=d["epochs"]; print("conv alt=$tv ppi", e["cluster_gcn_ppi_2bit"]["grouped_graph"]["epoch_ms"], "gin", e["batched_gin_2bit"]["epoch_ms"], "arxiv", [e["cluster_gcn_arxiv"][b]["epoch_ms"] for b in ["1_bit","2_bit","4_bit","8_bit"]])
PY
done
