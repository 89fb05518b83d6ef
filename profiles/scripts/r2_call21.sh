#!/bin/bash
O=gpurun_out
: > $O/r2c21_ab.txt
for c in 7 5 4; do
QGTC_TC05_FIXED=0 QGTC_TC05_CPS4=$c timeout 300 python profiles/scripts/tc05_ab.py fixed0_cps4_$c quick 2>&1 | grep '"bits": [12],' >> $O/r2c21_ab.txt
done
cat $O/r2c21_ab.txt
