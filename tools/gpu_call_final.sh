#!/bin/bash
# Runs on the GPU box: the last check of the round -- all GPU tests, smoke, and the default / config-3 bench lines.
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/gpu_tests_final.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_final.log 2>&1
timeout 200 python bench.py --steps 50 --warmup 5 > $O/bench_final.json 2> $O/bench_final.err
timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 > $O/bench_cfg3_final.json 2>> $O/bench_final.err
timeout 200 python bench.py --workload tick --steps 10 --warmup 3 > $O/bench_tick_final.json 2>> $O/bench_final.err
cat $O/gpu_tests_final.log $O/smoke_final.log; cut -c1-200 $O/bench_final.json; cut -c1-200 $O/bench_cfg3_final.json; cut -c1-200 $O/bench_tick_final.json; tail -2 $O/bench_final.err
