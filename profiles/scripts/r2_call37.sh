#!/bin/bash
# compute-sanitizer over the reduced parity set with the round-2 issuer protocol (owner-only commits, per-owner full barriers, single epilogue poller)
O=gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python profiles/scripts/sanitize_subset.py > $O/r2c37_sanitizer_$tool.txt 2>&1
  echo "exit $?" >> $O/r2c37_sanitizer_$tool.txt
  tail -5 $O/r2c37_sanitizer_$tool.txt
done
