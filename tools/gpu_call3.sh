#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'swing_kernel|plan_kernel|adapt_kernel|torque_cmd_kernel' -f -o $O/prof_aux \
    python tools/prof_aux.py > $O/prof_aux.log 2>&1
ncu -i $O/prof_aux.ncu-rep --page raw --csv > $O/prof_aux_raw.csv 2>/dev/null
tail -3 $O/prof_aux.log
