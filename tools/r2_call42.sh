#!/bin/bash
# Round 2, GPU call 42: compile-time variants against the in-tree build (tools/time_lib.py): set-up pass at 3 CTAs per SM
# (168 registers), loop-kernel refill threshold 1 / 4 / 8 idle lanes (default 2).
O=gpurun_out
mkdir -p $O
{ timeout 100 python tools/time_lib.py
  for L in scratch/libs/*.so; do QPB_LIB=$PWD/$L timeout 100 python tools/time_lib.py; done
  timeout 100 python tools/time_lib.py; } 2> $O/r2c42_variants.err | tee $O/r2c42_variants.txt
tail -3 $O/r2c42_variants.err
