#!/bin/bash
# call 68 (second build of the same change, no spills): grouped engine without run-time divisions in the prologue / role preambles and with the stage re-use scan only when some
# problems share x2: parity tests, fuzz, epochs (QGTC_GRP_REUSE_SCAN=1 = always scan), fresh start-up trace
O=gpurun_out
timeout 900 python -m pytest tests/test_grp_gpu.py tests/test_drivers_gpu.py tests/test_reference_gpu.py -m gpu -q -x 2>&1 | tail -3 > $O/r2c68_tests.txt
cat $O/r2c68_tests.txt
timeout 200 python profiles/scripts/fuzz_grp.py 40 5 2>&1 | tail -3
for tv in 1 0 0; do
QGTC_GRP_REUSE_SCAN=$tv timeout 600 python bench.py --steps 3 --warmup 3 --no-zero-tile --no-cpu-baseline > $O/r2c68_bench_t$tv.json 2> $O/r2c68_bench_t$tv.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c68_bench_t$tv.json").read().strip().splitlines()[-1])
e=d["epochs"]; print("always scan=$tv ppi", e["cluster_gcn_ppi_2bit"]["grouped_graph"]["epoch_ms"], "gin", e["batched_gin_2bit"]["epoch_ms"], "arxiv", [e["cluster_gcn_arxiv"][b]["epoch_ms"] for b in ["1_bit","2_bit","4_bit","8_bit"]])
PY
done
QGTC_GRP_TRACE=1 timeout 200 python profiles/scripts/grp_trace.py agg 759 75 2 16 > $O/r2c68_grp_trace_ppi_agg.txt 2>&1
sed -n 1,22p $O/r2c68_grp_trace_ppi_agg.txt
