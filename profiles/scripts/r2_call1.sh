#!/bin/bash
# Round-2 GPU call 1: parity (new worst-case-sum / grid / arxiv-chain tests), compute-sanitizer subset, ncu of the
# pack / prep kernels, cuBLASLt INT8 at the Fig. 8a shapes.  Everything lands in gpurun_out/.
O=gpurun_out
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/r2_pytest_gpu.txt
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python profiles/scripts/sanitize_subset.py > $O/r2_sanitizer_$tool.txt 2>&1
  echo "exit $?" >> $O/r2_sanitizer_$tool.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pack_row|pack_col" -s 4 -c 12 -o $O/r2_prof_pack -f \
    python profiles/scripts/pack_bench.py > $O/r2_pack_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"prep_b" -s 8 -c 4 -o $O/r2_prof_prep -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-epoch > /dev/null 2>&1
timeout 600 python profiles/scripts/int8_fig8a.py > $O/r2_int8_fig8a.txt 2>&1
timeout 300 python profiles/scripts/pack_bench.py > $O/r2_pack_bench.txt 2>&1
cat $O/r2_smoke.txt; cat $O/r2_pytest_gpu.txt; tail -3 $O/r2_sanitizer_*.txt; cat $O/r2_int8_fig8a.txt
