#!/bin/bash
# Round 2, GPU call 44: the state of the tree at the end of the round -- every GPU test, smoke(), the default bench line,
# the reference arm, config-3 line.
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4 > $O/r2c44_gpu_tests.log; cat $O/r2c44_gpu_tests.log
timeout 150 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c44_smoke.log 2>&1; tail -3 $O/r2c44_smoke.log
timeout 250 python bench.py --steps 50 --warmup 5 > $O/r2c44_bench.json 2> $O/r2c44_bench.err
timeout 250 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2c44_bench_reference.json 2>> $O/r2c44_bench.err
timeout 250 python bench.py --workload cfg3 --steps 10 --warmup 3 > $O/r2c44_bench_cfg3.json 2>> $O/r2c44_bench.err
for f in r2c44_bench r2c44_bench_reference r2c44_bench_cfg3; do cut -c1-260 $O/$f.json; done; tail -2 $O/r2c44_bench.err
