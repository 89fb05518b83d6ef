#!/bin/bash
# 2-GPU validation of the multi-rank paths (bench both arms, driver with logits gather) + pack bench (1 GPU part first)
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x 2>&1 | tail -3
QGTC_PACK_ROW=word timeout 300 python profiles/scripts/pack_bench.py > $O/r2c10_pack_bench_word.txt 2>&1
grep '"col_major": false' $O/r2c10_pack_bench_word.txt | head -6 | cut -c1-200
timeout 300 python profiles/scripts/pack_bench.py > $O/r2c10_pack_bench_auto.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2c10_bench_n2.json 2> $O/r2c10_bench_n2.err
tail -c 2500 $O/r2c10_bench_n2.json; tail -5 $O/r2c10_bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/r2c10_bench_ref_n2.json 2> $O/r2c10_bench_ref_n2.err
tail -c 1200 $O/r2c10_bench_ref_n2.json; tail -5 $O/r2c10_bench_ref_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 -m qgtc_ppopp22_b200.main_qgtc --dataset ppi --use_QGTC --n-classes 121 --grouped --cuda-graph --min-region-ms 100 --gather-logits > $O/r2c10_driver_n2.txt 2>&1
tail -4 $O/r2c10_driver_n2.txt
