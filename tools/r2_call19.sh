#!/bin/bash
# Round 2, GPU call 19: end-to-end pipeline: stages per batch (with per-slot scratch instead of cudaMallocAsync).
O=gpurun_out
mkdir -p $O
for S in 8 4 16 6; do
QPB_HOST_STAGES=$S timeout 300 python bench.py --no-secondary --steps 20 --warmup 3 2>/dev/null > $O/r2c19_bench_s$S.json
python - <<PY
import json
d = json.load(open("gpurun_out/r2c19_bench_s$S.json")); e = d["e2e"]
print("stages=$S value %.3e e2e %.3e sync %.3e bound %.3e frac %.3f" % (d["value"], e["value"], e["sync_call_value"], e["pcie_bound_qps"], e["pcie_frac"]))
PY
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "async or three_entry or small_batches or multi_device or warm" 2>&1 | tail -3
