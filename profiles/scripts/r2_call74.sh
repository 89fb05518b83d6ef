#!/bin/bash
# final-build validation: smoke(), all GPU tests, bench both arms at N=1
O=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $O/r2c74_smoke.txt 2>&1; tail -3 $O/r2c74_smoke.txt
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > $O/r2c74_pytest_gpu.txt
cat $O/r2c74_pytest_gpu.txt
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2c74_bench.json 2> $O/r2c74_bench.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 2 > $O/r2c74_bench_ref.json 2> $O/r2c74_bench_ref.err
tail -3 $O/r2c74_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c74_bench.json").read().strip().splitlines()[-1])
for k in ["value","ms_per_step","per_bit_tops_per_call","per_bit_tops_gemm_kernel","profile_convention_tops","single_gemm_per_launch_tops","clocks","parity_spot_check","gpu_launches"]:
    print(k, d.get(k))
print("e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], d["roofline"]["achieved"], d["roofline"].get("frac_of_sustained_peak"))
e=d["epochs"]; print("ppi", e["cluster_gcn_ppi_2bit"]["grouped_graph"]["epoch_ms"], "gin", e["batched_gin_2bit"]["epoch_ms"], "arxiv", [e["cluster_gcn_arxiv"][b]["epoch_ms"] for b in ["1_bit","2_bit","4_bit","8_bit"]])
r=json.loads(open("gpurun_out/r2c74_bench_ref.json").read().strip().splitlines()[-1])
print("reference arm", r["value"], r.get("profile_convention_tops"), r["e2e"]["value"], r["epochs"])
PY
