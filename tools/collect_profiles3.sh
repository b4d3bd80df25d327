#!/bin/bash
# Runs on the GPU box (under gpurun): end-of-round evidence, most important first (the box budget may cut the tail).
# usage: tools/collect_profiles3.sh <tag>
TAG=${1:-r01f}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 > $O/gpu_tests_${TAG}.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_${TAG}.log 2>&1
timeout 200 python bench.py --steps 50 --warmup 5 > $O/bench_${TAG}.json 2> $O/bench_${TAG}.err
timeout 200 python bench.py --workload cfg4 --steps 10 --warmup 3 > $O/bench_cfg4_${TAG}.json 2>> $O/bench_${TAG}.err
timeout 200 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_ref_${TAG}.json 2>> $O/bench_${TAG}.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file $O/launches_cfg4_${TAG}.csv \
    python bench.py --workload cfg4 --steps 3 --warmup 3 > $O/bench_cfg4_under_ncu_${TAG}.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_${TAG}.csv \
    python bench.py --steps 10 --warmup 3 > $O/bench_under_ncu_${TAG}.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mpc_qp -s 2 -c 1 -f -o $O/prof_mpc_${TAG} \
    python tools/time_mpc.py 16384 mixed > $O/prof_mpc_${TAG}.log 2>&1
python tools/ncu_digest.py $O/prof_mpc_${TAG}.ncu-rep 16384 > $O/prof_mpc_${TAG}_digest.txt 2>&1
timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 > $O/bench_cfg3_${TAG}.json 2>> $O/bench_${TAG}.err
timeout 200 python bench.py --workload tick --steps 10 --warmup 3 > $O/bench_tick_${TAG}.json 2>> $O/bench_${TAG}.err
if [ -f quadruped_control_b200/libqpb200_prof.so ]; then
  export QPB_LIB=$PWD/quadruped_control_b200/libqpb200_prof.so
  timeout 100 python tools/time_mpc.py 16384 mixed > $O/mpc_phase_${TAG}.log 2>&1
  for g in stand trot crawl; do timeout 100 python tools/time_mpc.py 8192 $g >> $O/mpc_phase_${TAG}.log 2>&1; done
  unset QPB_LIB
fi
cat $O/gpu_tests_${TAG}.log $O/smoke_${TAG}.log; cut -c1-260 $O/bench_${TAG}.json; cut -c1-260 $O/bench_cfg4_${TAG}.json; cut -c1-200 $O/bench_ref_${TAG}.json; tail -3 $O/bench_${TAG}.err
head -12 $O/prof_mpc_${TAG}_digest.txt
