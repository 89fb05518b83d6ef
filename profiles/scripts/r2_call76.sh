#!/bin/bash
# 4-GPU validation of the final build (our arm)
O=gpurun_out
mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29661 bench.py --gpus 4 --steps 20 --warmup 5 > $O/r2c76_bench_n4.json 2> $O/r2c76_bench_n4.err
tail -3 $O/r2c76_bench_n4.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c76_bench_n4.json").read().strip().splitlines()[-1])
for k in ["value","n_gpus","ms_per_step","rank_ms_timed_region","clocks","parity_spot_check"]:
    print(k, d.get(k))
print("e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"])
e=d["epochs"]; print("ppi", e["cluster_gcn_ppi_2bit"]["grouped_graph"]["epoch_ms"], "gin", e["batched_gin_2bit"]["epoch_ms"], "arxiv", [e["cluster_gcn_arxiv"][b]["epoch_ms"] for b in ["1_bit","2_bit","4_bit","8_bit"]])
PY
