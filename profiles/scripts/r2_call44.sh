#!/bin/bash
O=gpurun_out
timeout 900 python -m pytest tests/test_grp_gpu.py tests/test_drivers_gpu.py tests/test_reference_gpu.py -m gpu -q -x 2>&1 | tail -3 > $O/r2c44_tests.txt
cat $O/r2c44_tests.txt
for pf in 0 1; do
QGTC_GRP_PREFETCH=$pf timeout 600 python bench.py --steps 3 --warmup 3 --no-zero-tile --no-cpu-baseline > $O/r2c44_bench_pf$pf.json 2> $O/r2c44_bench_pf$pf.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c44_bench_pf$pf.json").read().strip().splitlines()[-1])
e=d["epochs"]; print("prefetch=$pf ppi", e["cluster_gcn_ppi_2bit"]["grouped_graph"]["epoch_ms"], "gin", e["batched_gin_2bit"]["epoch_ms"], "arxiv", [e["cluster_gcn_arxiv"][b]["epoch_ms"] for b in ["1_bit","2_bit","4_bit","8_bit"]])
PY
done
QGTC_GRP_TRACE=1 timeout 200 python profiles/scripts/grp_trace.py agg 759 75 2 16 > $O/r2c44_grp_trace_ppi.txt 2>&1
sed -n 3,30p $O/r2c44_grp_trace_ppi.txt
