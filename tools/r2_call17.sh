#!/bin/bash
# Round 2, GPU call 17: lane-shared epilogue + completion words for the one-robot path: parity, shim latency, small batches.
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_sanitizer_gpu.py -m gpu -q -x 2>&1 | tail -25 | cut -c1-220 | sed "s/^/parity: /"
for Pn in 1 0; do QPB_SMALL_POLL=$Pn timeout 120 ./quadruped_control_b200/cpp/shim_latency 2>&1 | tail -1 | sed "s/^/poll=$Pn: /" | tee -a $O/r2c17_shim_latency.txt; done
timeout 300 python tools/time_small_batches.py 2>&1 | tail -9 | tee $O/r2c17_small_batches.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:tpq_one -s 2 -c 1 -f -o $O/r2c17_one_n1 python tools/prof_one.py 1 4 > $O/r2c17_one_n1.log 2>&1; tail -1 $O/r2c17_one_n1.log
