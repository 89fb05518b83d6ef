#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > $O/r2c8_pytest_gpu.txt
cat $O/r2c8_pytest_gpu.txt
if grep -q "failed\|error" $O/r2c8_pytest_gpu.txt; then exit 1; fi
timeout 300 python profiles/scripts/pack_bench.py > $O/r2c8_pack_bench.txt 2>&1
cat $O/r2c8_pack_bench.txt
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python profiles/scripts/sanitize_subset.py > $O/r2c8_sanitizer_$tool.txt 2>&1
  echo "exit $?" >> $O/r2c8_sanitizer_$tool.txt
  tail -4 $O/r2c8_sanitizer_$tool.txt
done
timeout 900 python bench.py --steps 20 --warmup 5 --no-epoch > $O/r2c8_bench.json 2> $O/r2c8_bench.err
python -c "
import json; d=json.load(open('$O/r2c8_bench.json')); print(json.dumps(d['zero_tile_sweep'])[:3000]); print(d['value'], d['e2e']['value'])"
tail -3 $O/r2c8_bench.err
