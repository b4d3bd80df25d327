#!/bin/bash
# Round 2, GPU call 40: finishing pass overlapped with the tail of the loop pass (programmatic dependent launch + queue of
# finished records) against three serial launches; the GPU tests that touch the balance path; launch list; bench line.
O=gpurun_out
mkdir -p $O
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "overlapped or graph or profiles or golden" 2>&1 | tail -5
timeout 300 python tools/time_overlap.py > $O/r2c40_overlap_ab.txt 2> $O/r2c40_overlap_ab.err
cat $O/r2c40_overlap_ab.txt; tail -3 $O/r2c40_overlap_ab.err
timeout 560 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=4 2>&1 | tail -12 > $O/r2c40_gpu_tests.log
cat $O/r2c40_gpu_tests.log
timeout 200 python bench.py --steps 50 --warmup 5 > $O/r2c40_bench.json 2> $O/r2c40_bench.err
cut -c1-300 $O/r2c40_bench.json; tail -2 $O/r2c40_bench.err
