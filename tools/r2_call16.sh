#!/bin/bash
# Round 2, GPU call 16: where a single warm QP spends its time (ncu source view of tpq_one_kernel at n = 1), and the warm
# tick at 65536 records: one-launch kernel against the three-pass path.
O=gpurun_out
mkdir -p $O
timeout 300 ncu --set full --import-source on --clock-control none -k regex:tpq_one -s 2 -c 1 -f -o $O/r2c16_one_n1 python tools/prof_one.py 1 4 > $O/r2c16_one_n1.log 2>&1; tail -2 $O/r2c16_one_n1.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:tpq_one -s 2 -c 1 -f -o $O/r2c16_one_n1_cold python tools/prof_one.py 1 4 cold > $O/r2c16_one_n1_cold.log 2>&1; tail -2 $O/r2c16_one_n1_cold.log
for M in 1099511627776 0; do
QPB_TPQ_WARM_ONE_MAX=$M timeout 400 python bench.py --steps 20 --warmup 3 2>/dev/null > $O/r2c16_bench_warm_$M.json
python - <<PY
import json
d = json.load(open("gpurun_out/r2c16_bench_warm_$M.json"))
w = d["secondary"]["cfg2_warm_tick"]
print("warm_one_max=$M", w["value"], w["ms_per_launch"], w["kernels_per_call"], w["iters_mean"])
PY
done
