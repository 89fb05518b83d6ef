#!/bin/bash
# 8 GPUs: end-to-end leg with and without binding each rank to its GPU's NUMA node
O=gpurun_out
nvidia-smi topo -m > $O/r2c40_topo.txt 2>&1; head -14 $O/r2c40_topo.txt | cut -c1-160
for numa in 0 1; do
QGTC_BENCH_NUMA=$numa timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2963$numa bench.py --gpus 8 --steps 5 --warmup 3 --no-epoch --no-zero-tile > $O/r2c40_bench_n8_numa$numa.json 2> $O/r2c40_bench_n8_numa$numa.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2c40_bench_n8_numa$numa.json").read().strip().splitlines()[-1])
print("numa=$numa value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["rank_s_per_step"], d["e2e"].get("numa_node_rank0"))
PY
done
