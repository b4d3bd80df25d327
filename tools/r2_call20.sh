#!/bin/bash
# Round 2, GPU call 20: wire records (488 B up / 200 B down): parity, end-to-end against the padded records; full suite.
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | cut -c1-200 | sed "s/^/suite: /"
for i in 1 2; do
timeout 300 python bench.py --no-secondary --steps 20 --warmup 3 2>/dev/null > $O/r2c20_bench_$i.json
python - <<PY
import json
d = json.load(open("gpurun_out/r2c20_bench_$i.json")); e = d["e2e"]
print("value %.3e e2e(wire) %.3e sync %.3e | padded %.3e sync %.3e | bound %.3e = %.1f GB/s frac %.3f launches %d" % (d["value"], e["value"], e["sync_call_value"], e["padded_records_value"], e["padded_records_sync_value"], e["pcie_bound_qps"], e["pcie_bound_gbs"], e["pcie_frac"], d["gpu_launches"]))
PY
done
