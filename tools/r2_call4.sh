#!/bin/bash
# Round 2, GPU call 4: three-kernel range-space path (set-up / loop / finish): parity, lanes-per-QP and occupancy A/B,
# launch list with per-kernel times.
O=gpurun_out
mkdir -p $O
QPB_TPQ_MIN_N=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "profiles or mask or bad_input or degenerate or warm or three_entry or config1 or general" 2>&1 | tail -3 | sed "s/^/minN0 lpq2: /"
for L in 4 1; do
  QPB_TPQ_MIN_N=0 QPB_TPQ_LPQ=$L timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "profiles or mask or bad_input or degenerate or warm" 2>&1 | tail -2 | sed "s/^/minN0 lpq$L: /"
done
for LIB in quadruped_control_b200/libqpb200.so scratch/libs/libqpb_occ_mid.so scratch/libs/libqpb_occ_hi.so; do
 for L in 1 2 4; do
  QPB_LIB=$PWD/$LIB QPB_TPQ_LPQ=$L timeout 200 python bench.py --steps 30 --warmup 5 2>/dev/null | cut -c1-120 | sed "s|^|$(basename $LIB) lpq$L cfg2: |"
  QPB_LIB=$PWD/$LIB QPB_TPQ_LPQ=$L timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 2>/dev/null | cut -c1-120 | sed "s|^|$(basename $LIB) lpq$L cfg3: |"
 done
done
for L in 2 4; do
QPB_TPQ_LPQ=$L timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:tpq_ -s 3 -c 3 --csv --log-file $O/r2c4_launches_cfg3_lpq$L.csv python tools/prof_run.py cfg3 3 > /dev/null 2>&1
echo "== launches lpq$L"; grep -E "tpq_" $O/r2c4_launches_cfg3_lpq$L.csv | awk -F'","' '{print $5, $(NF-2), $(NF)}' | cut -c1-200
done
