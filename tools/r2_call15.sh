#!/bin/bash
# Round 2, GPU call 15: full GPU suite; working-set word from every kernel, warm batches on the one-launch kernel;
# launch + wait floor; shim latency; bench with the warm-tick secondary entry.
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 | cut -c1-220 | sed "s/^/suite: /"
timeout 60 ./scratch/ubench_launch | tee $O/r2c15_launch_floor.txt
timeout 120 ./quadruped_control_b200/cpp/shim_latency 2>&1 | tail -1 | tee $O/r2c15_shim_latency.txt
timeout 400 python bench.py --steps 20 --warmup 3 2>/dev/null | tee $O/r2c15_bench_cfg2.json | cut -c1-150
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c15_bench_cfg2.json"))
print(json.dumps(d["secondary"].get("cfg2_warm_tick"))[:600])
print("e2e", d["e2e"]["value"], d["e2e"].get("sync_call_value"))
PY
