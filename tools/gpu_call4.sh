#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_plan.py tests/test_swing.py -m gpu -q -x 2>&1 | tail -8 > $O/gpu_tests_c4.log
timeout 200 python tools/time_plan.py > $O/plan_kernels_c4.txt 2>&1
timeout 200 python bench.py --workload tick --steps 10 --warmup 3 > $O/bench_tick_c4.json 2> $O/bench_c4.err
cat $O/gpu_tests_c4.log $O/plan_kernels_c4.txt; cut -c1-330 $O/bench_tick_c4.json; tail -3 $O/bench_c4.err
