#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > $O/r2c9_pytest_gpu.txt
cat $O/r2c9_pytest_gpu.txt
if grep -q "failed\|error" $O/r2c9_pytest_gpu.txt; then exit 1; fi
timeout 600 python profiles/scripts/epoch_modes.py ppi 2 > $O/r2c9_epoch_ppi.txt 2>&1
timeout 600 python profiles/scripts/epoch_modes.py ogbn-arxiv 2 8 > $O/r2c9_epoch_arxiv.txt 2>&1
timeout 600 python profiles/scripts/epoch_modes.py gin 2 > $O/r2c9_epoch_gin.txt 2>&1
cat $O/r2c9_epoch_*.txt
timeout 120 python profiles/scripts/grp_trace.py xw > $O/r2c9_trace_xw.txt 2>&1
head -12 $O/r2c9_trace_xw.txt
for mode in shfl word; do
  QGTC_PACK_ROW=$mode timeout 300 python profiles/scripts/pack_bench.py > $O/r2c9_pack_bench_$mode.txt 2>&1
  echo "== $mode"; grep '"col_major": false' $O/r2c9_pack_bench_$mode.txt | cut -c1-200
done
QGTC_PACK_ROW=word timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pack_row_kernel" -s 4 -c 3 -o $O/r2c9_prof_pack_word1 -f \
    python profiles/scripts/pack_bench.py > /dev/null 2>&1
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python profiles/scripts/sanitize_subset.py > $O/r2c9_sanitizer_racecheck.txt 2>&1
echo "exit $?" >> $O/r2c9_sanitizer_racecheck.txt
tail -4 $O/r2c9_sanitizer_racecheck.txt
