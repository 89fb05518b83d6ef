#!/bin/bash
# 8-GPU validation of both bench arms exactly as the driver launches them
O=gpurun_out
mkdir -p $O
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r2c14_bench_n8.json 2> $O/r2c14_bench_n8.err
tail -c 1200 $O/r2c14_bench_n8.json; tail -3 $O/r2c14_bench_n8.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29712 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > $O/r2c14_bench_ref_n8.json 2> $O/r2c14_bench_ref_n8.err
tail -c 800 $O/r2c14_bench_ref_n8.json; tail -3 $O/r2c14_bench_ref_n8.err
