#!/bin/bash
# A/B of experiment builds of the library (scratch/*.so) on the packed balance kernel.
O=gpurun_out
mkdir -p $O
: > $O/ab.txt
for lib in quadruped_control_b200/libqpb200.so scratch/*.so; do
  for wl in cfg2 cfg3; do
    QPB_LIB=$PWD/$lib timeout 120 python tests/tools/time_kernel.py $wl >> $O/ab.txt 2>&1
  done
done
cat $O/ab.txt
