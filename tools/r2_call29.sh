#!/bin/bash
# Round 2, GPU call 29: warm batches: early finish in the set-up pass + worklist finish, against the one-launch kernel.
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_sanitizer_gpu.py -m gpu -q -x 2>&1 | tail -12 | cut -c1-250 | sed "s/^/parity: /"
for M in 16384 1099511627776; do
for i in 1 2; do
QPB_TPQ_WARM_DEFER_MIN=$M timeout 400 python bench.py --steps 20 --warmup 3 2>/dev/null > $O/r2c29_bench_$M.json
python - <<PY
import json
d = json.load(open("gpurun_out/r2c29_bench_$M.json"))
w = d["secondary"]["cfg2_warm_tick"]
print("defer_min=$M warm tick %.3e  %.1f us  kernels %d  iters %.3f  err %.1e | cold value %.3e" % (w["value"], w["ms_per_launch"] * 1e3, w["kernels_per_call"], w["iters_mean"], w["max_rel_grf_err_vs_oracle"], d["value"]))
PY
done
done
