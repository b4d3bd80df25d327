#!/bin/bash
# Round 2, GPU call 6: rotated loop + worklist compaction: parity (all sizes through the range-space path), bench, per-kernel times.
O=gpurun_out
mkdir -p $O
QPB_TPQ_MIN_N=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "profiles or mask or bad_input or degenerate or warm or three_entry or general or mappings or kkt" 2>&1 | tail -3 | sed "s/^/minN0 lpq2: /"
for L in 1 4; do QPB_TPQ_LPQ=$L QPB_TPQ_MIN_N=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "profiles or mask or bad_input or degenerate or warm or mappings" 2>&1 | tail -2 | sed "s/^/minN0 lpq$L: /"; done
for L in 1 2 4; do
  QPB_TPQ_LPQ=$L timeout 200 python bench.py --steps 30 --warmup 5 2>/dev/null | cut -c1-120 | sed "s|^|lpq$L cfg2: |"
  QPB_TPQ_LPQ=$L timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 2>/dev/null | cut -c1-120 | sed "s|^|lpq$L cfg3: |"
done
for W in cfg2 cfg3; do
QPB_TPQ_LPQ=1 timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:tpq_ -s 3 -c 3 --csv --log-file $O/r2c6_launches_${W}_lpq1.csv python tools/prof_run.py $W 3 > /dev/null 2>&1
echo "== launches $W lpq1"; grep -E "tpq_" $O/r2c6_launches_${W}_lpq1.csv | awk -F'","' '{print substr($5,1,40), $(NF-2), $(NF)}' | cut -c1-160
done
