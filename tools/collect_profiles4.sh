#!/bin/bash
# Runs on the GPU box (under gpurun): end-of-round evidence with the final library, most important first.
TAG=${1:-r01g}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/gpu_tests_${TAG}.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_${TAG}.log 2>&1
timeout 200 python bench.py --steps 50 --warmup 5 > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err
timeout 200 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_ref_${TAG}.json 2>> $O/bench_${TAG}.err
timeout 200 python bench.py --profile light --steps 50 --warmup 5 > $O/bench_light_${TAG}.json 2>> $O/bench_${TAG}.err
timeout 200 python bench.py --profile stress --steps 50 --warmup 5 > $O/bench_stress_${TAG}.json 2>> $O/bench_${TAG}.err
timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 > $O/bench_cfg3_${TAG}.json 2>> $O/bench_${TAG}.err
timeout 200 python bench.py --workload tick --steps 10 --warmup 3 > $O/bench_tick_${TAG}.json 2>> $O/bench_${TAG}.err
timeout 200 python bench.py --workload cfg4 --steps 10 --warmup 3 > $O/bench_cfg4_${TAG}.json 2>> $O/bench_${TAG}.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_${TAG}.csv \
    python bench.py --steps 10 --warmup 3 > $O/bench_under_ncu_${TAG}.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_tick_${TAG}.csv \
    python bench.py --workload tick --steps 3 --warmup 3 > $O/bench_tick_under_ncu_${TAG}.log 2>&1
timeout 100 python tools/time_plan.py > $O/plan_kernels_${TAG}.txt 2>&1
cat $O/gpu_tests_${TAG}.log $O/smoke_${TAG}.log; for f in bench bench_ref bench_light bench_stress bench_cfg3 bench_tick bench_cfg4; do cut -c1-200 $O/${f}_${TAG}.json; done; tail -3 $O/bench_${TAG}.err
