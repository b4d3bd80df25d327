#!/bin/bash
# Round 2, GPU call 10: multipliers / lever arms in shared memory, loop kernel at 12 (and 16) warps per SM.
O=gpurun_out
mkdir -p $O
QPB_TPQ_MIN_N=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "profiles or mask or bad_input or degenerate or warm or three_entry or general or mappings or kkt" 2>&1 | tail -3 | sed "s/^/minN0 lpq2: /"
for L in 1 4; do QPB_TPQ_LPQ=$L QPB_TPQ_MIN_N=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "profiles or mask or bad_input or degenerate or warm or mappings" 2>&1 | tail -2 | sed "s/^/minN0 lpq$L: /"; done
for LIB in quadruped_control_b200/libqpb200.so scratch/libs/libqpb_r128.so scratch/libs/libqpb_r255.so; do
for L in 1 2 4; do
  QPB_LIB=$PWD/$LIB QPB_TPQ_LPQ=$L timeout 200 python bench.py --no-secondary --steps 30 --warmup 5 2>/dev/null | cut -c1-120 | sed "s|^|$(basename $LIB) lpq$L cfg2: |"
  QPB_LIB=$PWD/$LIB QPB_TPQ_LPQ=$L timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 2>/dev/null | cut -c1-120 | sed "s|^|$(basename $LIB) lpq$L cfg3: |"
done
done
for W in cfg2 cfg3; do
QPB_TPQ_LPQ=1 timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:tpq_loop -s 1 -c 1 --csv --log-file $O/r2c10_launches_${W}_lpq1.csv python tools/prof_run.py $W 3 > /dev/null 2>&1
echo "== loop $W lpq1"; grep -E "tpq_" $O/r2c10_launches_${W}_lpq1.csv | awk -F'","' '{print substr($5,1,40), $(NF-2), $(NF)}' | cut -c1-160
done
