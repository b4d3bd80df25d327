#!/bin/bash
# Round 2, first GPU call: smoke (short timeout, guards against a hung kernel), GPU tests, A/B bench of the
# thread-per-QP kernel against the half-warp kernel on the same box.
O=gpurun_out
mkdir -p $O
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c1_smoke.log 2>&1; echo "smoke rc=$?" >> $O/r2c1_smoke.log
tail -3 $O/r2c1_smoke.log
if ! grep -q "smoke rc=0" $O/r2c1_smoke.log; then exit 1; fi
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/r2c1_gpu_tests.log; cat $O/r2c1_gpu_tests.log
timeout 200 python bench.py --steps 50 --warmup 5 > $O/r2c1_bench_tpq.json 2> $O/r2c1_bench.err
QPB_QPS_PER_WARP=2 timeout 200 python bench.py --steps 50 --warmup 5 > $O/r2c1_bench_k16.json 2>> $O/r2c1_bench.err
timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 > $O/r2c1_bench_cfg3_tpq.json 2>> $O/r2c1_bench.err
QPB_QPS_PER_WARP=2 timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 > $O/r2c1_bench_cfg3_k16.json 2>> $O/r2c1_bench.err
for f in $O/r2c1_bench_*.json; do echo $f; cut -c1-400 $f; done; tail -3 $O/r2c1_bench.err
