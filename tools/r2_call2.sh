#!/bin/bash
# Round 2, GPU call 2: remaining GPU tests, ncu full capture of the thread-per-QP kernel on cfg3, refill-threshold A/B.
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/r2c2_gpu_tests.log; cat $O/r2c2_gpu_tests.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:balance_qp_tpq -s 1 -c 1 -f -o $O/r2c2_prof_cfg3 \
    python tools/prof_run.py cfg3 3 > $O/r2c2_prof.log 2>&1
ncu -i $O/r2c2_prof_cfg3.ncu-rep --page raw --csv > $O/r2c2_prof_cfg3_raw.csv 2>/dev/null
python tools/ncu_digest.py $O/r2c2_prof_cfg3.ncu-rep 1048576 > $O/r2c2_prof_cfg3_digest.txt 2>&1
head -40 $O/r2c2_prof_cfg3_digest.txt
for R in 1 4 8; do
  QPB_LIB=$PWD/scratch/libs/libqpb_refill$R.so timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 2>/dev/null | cut -c1-160 | sed "s/^/refill$R cfg3: /"
  QPB_LIB=$PWD/scratch/libs/libqpb_refill$R.so timeout 200 python bench.py --steps 30 --warmup 5 2>/dev/null | cut -c1-160 | sed "s/^/refill$R cfg2: /"
done
