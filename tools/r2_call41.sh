#!/bin/bash
# Round 2, GPU call 41: per-CTA timeline of the loop / finishing passes (developer build, see tools/timeline.py).
O=gpurun_out
mkdir -p $O
QPB_LIB=$PWD/scratch/libs/libqpb_timeline.so timeout 200 python tools/timeline.py cfg2 > $O/r2c41_timeline_cfg2.txt 2> $O/r2c41_timeline.err
cat $O/r2c41_timeline_cfg2.txt; tail -3 $O/r2c41_timeline.err
QPB_LIB=$PWD/scratch/libs/libqpb_timeline.so timeout 200 python tools/timeline.py cfg3 > $O/r2c41_timeline_cfg3.txt 2>> $O/r2c41_timeline.err
cat $O/r2c41_timeline_cfg3.txt
