#!/bin/bash
# Round 2, GPU call 11: full GPU test suite (incl. compute-sanitizer and the direct reference-source comparison), bench, per-kernel times.
O=gpurun_out
mkdir -p $O
( time timeout 1700 python -m pytest tests -m gpu -q -x --durations=8 ) 2>&1 | tail -22
timeout 200 python bench.py --no-secondary --steps 30 --warmup 5 2>/dev/null | cut -c1-120 | sed "s|^|cfg2: |"
timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 2>/dev/null | cut -c1-120 | sed "s|^|cfg3: |"
for W in cfg2 cfg3; do
timeout 200 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:tpq_ -s 3 -c 3 --csv --log-file $O/r2c11_launches_${W}.csv python tools/prof_run.py $W 3 > /dev/null 2>&1
echo "== launches $W"; grep -E "tpq_" $O/r2c11_launches_${W}.csv | awk -F'","' '{print substr($5,1,40), $(NF-2), $(NF)}' | cut -c1-160
done
