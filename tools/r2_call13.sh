#!/bin/bash
# Round 2, GPU call 13: longest-first worklist sections: parity, bench, loop kernel times.
O=gpurun_out
mkdir -p $O
QPB_TPQ_MIN_N=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "profiles or mask or bad_input or degenerate or warm or three_entry or general or mappings or kkt or reference_sources or async" 2>&1 | tail -3 | sed "s/^/minN0: /"
for L in 2 4; do QPB_TPQ_LPQ=$L QPB_TPQ_MIN_N=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "profiles or mask or bad_input or warm or mappings" 2>&1 | tail -2 | sed "s/^/minN0 lpq$L: /"; done
for i in 1 2; do
timeout 200 python bench.py --no-secondary --steps 30 --warmup 5 2>/dev/null | cut -c1-120 | sed "s|^|cfg2: |"
timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 2>/dev/null | cut -c1-120 | sed "s|^|cfg3: |"
done
for W in cfg2 cfg3; do
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:tpq_loop -s 1 -c 1 --csv --log-file $O/r2c13_launches_${W}.csv python tools/prof_run.py $W 3 > /dev/null 2>&1
echo "== loop $W"; grep -E "tpq_" $O/r2c13_launches_${W}.csv | awk -F'","' '{print substr($5,1,40), $(NF-2), $(NF)}' | cut -c1-160
done
