#!/bin/bash
O=gpurun_out
mkdir -p $O
QGTC_TC05_TRACE=1 QGTC_TC05_TRACE_F4=1 timeout 120 python profiles/scripts/tc05_trace.py 64 > $O/r2c18_trace_f4_n64.txt 2>&1
QGTC_TC05_TRACE=1 QGTC_TC05_TRACE_F4=1 timeout 120 python profiles/scripts/tc05_trace.py 16 > $O/r2c18_trace_f4_n16.txt 2>&1
QGTC_TC05_TRACE=1 timeout 120 python profiles/scripts/tc05_trace.py 64 > $O/r2c18_trace_u8_n64.txt 2>&1
cat $O/r2c18_trace_f4_n64.txt
: > $O/r2c18_ab.txt
timeout 300 python profiles/scripts/tc05_ab.py fixed1_v2 quick >> $O/r2c18_ab.txt 2>&1
cat $O/r2c18_ab.txt
