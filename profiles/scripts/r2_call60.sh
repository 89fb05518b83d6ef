#!/bin/bash
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_grp -s 4 -c 4 -o $O/r2c60_prof_grp_ppi -f \
    python profiles/scripts/grouped_epoch_once.py ppi 2 2 > $O/r2c60_prof.log 2>&1
tail -3 $O/r2c60_prof.log; ls -la $O/r2c60_prof_grp_ppi.ncu-rep
