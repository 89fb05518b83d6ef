#!/bin/bash
# experiment: grouped kernel with its run-time switches collapsed to the 2-bit case (code size 24.6K -> ~10K SASS lines per instance): instruction-cache hypothesis
O=gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-zero-tile --no-cpu-baseline > $O/r2c55_bench.json 2> $O/r2c55_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c55_bench.json").read().strip().splitlines()[-1])
e=d["epochs"]; print("HACK2 ppi(2-bit)", e["cluster_gcn_ppi_2bit"]["grouped_graph"]["epoch_ms"], "gin(2-bit)", e["batched_gin_2bit"]["epoch_ms"], "arxiv 2-bit", e["cluster_gcn_arxiv"]["2_bit"]["epoch_ms"])
PY
