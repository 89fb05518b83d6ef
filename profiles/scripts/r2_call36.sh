#!/bin/bash
O=gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 --no-zero-tile > $O/r2c36_bench.json 2> $O/r2c36_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c36_bench.json").read().strip().splitlines()[-1])
for k in ["value","ms_per_step","per_bit_tops_per_call","per_bit_tops_gemm_kernel","profile_convention_tops","single_gemm_per_launch_tops","clocks","parity_spot_check"]:
    print(k, d.get(k))
print("e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], d["roofline"]["achieved"])
e=d["epochs"]; print("ppi", e["cluster_gcn_ppi_2bit"]["grouped_graph"]["epoch_ms"], "gin", e["batched_gin_2bit"]["epoch_ms"], "arxiv", [e["cluster_gcn_arxiv"][b]["epoch_ms"] for b in ["1_bit","2_bit","4_bit","8_bit"]])
PY
nvidia-smi --query-gpu=power.limit,power.draw,clocks.sm,clocks.max.sm,temperature.gpu --format=csv
