#!/bin/bash
# call 72: single-call latency sweep on the final build (grouped engine now launches with the 2.5 KB parameter block)
O=gpurun_out
timeout 300 python profiles/scripts/single_call_sweep.py > $O/r2c72_single_call.txt 2>&1
cat $O/r2c72_single_call.txt
