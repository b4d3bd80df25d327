#!/bin/bash
# Round 2, GPU call 47: every GPU test (with the full-size config-5 test) and the default bench line of the final tree.
O=gpurun_out
mkdir -p $O
timeout 500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=5 2>&1 | tail -12 > $O/r2c47_gpu_tests.log; cat $O/r2c47_gpu_tests.log
timeout 200 python bench.py --steps 50 --warmup 5 > $O/r2c47_bench.json 2> $O/r2c47_bench.err; cut -c1-200 $O/r2c47_bench.json; tail -2 $O/r2c47_bench.err
