#!/bin/bash
# Runs on the GPU box: tests + timings of the 256-bit record kernels and the single-process multi-device calls.
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/gpu_tests_c2.log
timeout 200 python tools/time_plan.py > $O/plan_kernels_c2.txt 2>&1
timeout 200 python tools/time_multi.py > $O/multi_c2.txt 2>&1
timeout 200 python bench.py --workload tick --steps 10 --warmup 3 > $O/bench_tick_c2.json 2> $O/bench_c2.err
cat $O/gpu_tests_c2.log $O/plan_kernels_c2.txt $O/multi_c2.txt; cut -c1-330 $O/bench_tick_c2.json; tail -3 $O/bench_c2.err
