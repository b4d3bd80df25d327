#!/bin/bash
# Round 2, GPU call 3: lanes-per-QP variants (1, 2, 4) of the range-space kernel: parity subset, bench A/B, ncu for 2 and 4.
O=gpurun_out
mkdir -p $O
for L in 2 4 1; do
  QPB_TPQ_LPQ=$L timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "profiles or mask or bad_input or degenerate or warm or three_entry or config1" 2>&1 | tail -3 | sed "s/^/lpq$L: /"
done
for L in 1 2 4; do
  QPB_TPQ_LPQ=$L timeout 200 python bench.py --steps 30 --warmup 5 2>/dev/null | cut -c1-150 | sed "s/^/lpq$L cfg2: /"
  QPB_TPQ_LPQ=$L timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 2>/dev/null | cut -c1-150 | sed "s/^/lpq$L cfg3: /"
done
for L in 2 4; do
  QPB_TPQ_LPQ=$L timeout 300 ncu --set full --clock-control none --import-source on -k regex:balance_qp_tpq -s 1 -c 1 -f -o $O/r2c3_prof_cfg3_lpq$L \
      python tools/prof_run.py cfg3 3 > $O/r2c3_prof.log 2>&1
  ncu -i $O/r2c3_prof_cfg3_lpq$L.ncu-rep --page raw --csv > $O/r2c3_prof_cfg3_lpq${L}_raw.csv 2>/dev/null
  python tools/ncu_digest.py $O/r2c3_prof_cfg3_lpq$L.ncu-rep 1048576 > $O/r2c3_prof_cfg3_lpq${L}_digest.txt 2>&1
  echo "== lpq$L"; head -34 $O/r2c3_prof_cfg3_lpq${L}_digest.txt | grep -E "duration|issue_active|fp64|inst_executed.sum|registers|warps_active|stalled|thread_inst|occupancy"
done
