#!/bin/bash
# Runs on the GPU box (under gpurun): bench line, ncu launch list of the same command, one full capture.
# usage: tools/collect_profiles.sh <tag>
TAG=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_${TAG}.json 2>> gpurun_out/bench_${TAG}.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 10 --warmup 3 > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:balance_qp -s 3 -c 1 -f -o gpurun_out/prof_${TAG} \
    python tools/prof_run.py cfg2 6 > gpurun_out/prof_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:balance_qp -s 1 -c 1 -f -o gpurun_out/prof_cfg3_${TAG} \
    python tools/prof_run.py cfg3 3 >> gpurun_out/prof_${TAG}.log 2>&1
cat gpurun_out/bench_${TAG}.json; cat gpurun_out/bench_ref_${TAG}.json | cut -c1-400; tail -3 gpurun_out/bench_${TAG}.err
