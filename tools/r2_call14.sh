#!/bin/bash
# Round 2, GPU call 14: full GPU suite on the current tree; one-launch range-space kernel (small batches): parity, latency
# per batch size (cold / warm), shim latency with and without it; finish kernel at 3 CTAs per SM (A/B).
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | sed "s/^/suite: /"
timeout 300 python tools/time_small_batches.py 2>&1 | tail -9 | tee $O/r2c14_small_batches.txt
for M in 0 128; do QPB_TPQ_ONE_MAX=$M timeout 120 ./quadruped_control_b200/cpp/shim_latency 2>&1 | tail -1 | sed "s/^/shim one_max=$M: /" | tee -a $O/r2c14_shim_latency.txt; done
for i in 1 2; do
for LIB in quadruped_control_b200/libqpb200.so scratch/libs/libqpb_f3.so; do
  QPB_LIB=$PWD/$LIB timeout 200 python bench.py --no-secondary --steps 30 --warmup 5 2>/dev/null | cut -c1-100 | sed "s|^|$(basename $LIB) cfg2: |"
  QPB_LIB=$PWD/$LIB timeout 200 python bench.py --no-secondary --workload cfg3 --steps 10 --warmup 3 2>/dev/null | cut -c1-100 | sed "s|^|$(basename $LIB) cfg3: |"
done
done
