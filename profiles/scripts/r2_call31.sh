#!/bin/bash
# evidence run after the issuer-protocol / expander changes of the tcgen05 engine: all GPU tests, ncu capture + launch list of the bench
# kernels, N=1 bench lines of both arms
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > $O/r2c31_pytest_gpu.txt
cat $O/r2c31_pytest_gpu.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc05 -s 12 -c 4 -o $O/r2c31_prof_tc05_bench -f \
    python bench.py --steps 1 --warmup 3 --passes 1 --no-cpu-baseline --no-epoch --no-zero-tile > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm_|prep_b" -s 24 -c 64 --csv --log-file $O/r2c31_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --passes 1 --no-cpu-baseline --no-epoch --no-zero-tile > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2c31_launches_epoch_ppi.csv \
    python profiles/scripts/grouped_epoch_once.py ppi 2 2 > $O/r2c31_launches_epoch_ppi.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2c31_bench.json 2> $O/r2c31_bench.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > $O/r2c31_bench_ref.json 2> $O/r2c31_bench_ref.err
tail -c 2500 $O/r2c31_bench.json | head -c 1200; echo; tail -3 $O/r2c31_bench.err; tail -c 600 $O/r2c31_bench_ref.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c31_bench.json").read().strip().splitlines()[-1])
for k in ["value","ms_per_step","per_bit_tops_per_call","per_bit_tops_gemm_kernel","profile_convention_tops","single_gemm_per_launch_tops","clocks","parity_spot_check"]:
    print(k, d.get(k))
print("e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], d["roofline"]["achieved"])
e=d["epochs"]; print("ppi", e["cluster_gcn_ppi_2bit"]["grouped_graph"]["epoch_ms"], "gin", e["batched_gin_2bit"]["epoch_ms"], "arxiv", [e["cluster_gcn_arxiv"][b]["epoch_ms"] for b in ["1_bit","2_bit","4_bit","8_bit"]])
PY
