#!/bin/bash
# call 66: finer clock64 trace of the grouped engine's start-up (set-up phases, load issue vs landing), ppi-shaped aggregation and X.W
O=gpurun_out
QGTC_GRP_TRACE=1 timeout 200 python profiles/scripts/grp_trace.py agg 759 75 2 16 > $O/r2c66_grp_trace_ppi_agg.txt 2>&1
QGTC_GRP_TRACE=1 timeout 200 python profiles/scripts/grp_trace.py xw 759 75 2 16 > $O/r2c66_grp_trace_ppi_xw.txt 2>&1
sed -n 1,45p $O/r2c66_grp_trace_ppi_agg.txt
