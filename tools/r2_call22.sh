#!/bin/bash
# Round 2, GPU call 22: compact by-value kernel parameters: host pipeline timeline, end to end, parity, shim latency.
O=gpurun_out
mkdir -p $O
timeout 120 python tools/trace_host.py wire 2> $O/r2c22_trace_wire.txt; tail -34 $O/r2c22_trace_wire.txt | head -14
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
for S in 8 12; do
QPB_HOST_STAGES=$S timeout 300 python bench.py --no-secondary --steps 20 --warmup 3 2>/dev/null > $O/r2c22_bench_$S.json
python - <<PY
import json
d = json.load(open("gpurun_out/r2c22_bench_$S.json")); e = d["e2e"]
print("stages=$S value %.3e e2e(wire) %.3e sync %.3e | padded %.3e sync %.3e | bound %.3e = %.1f GB/s frac %.3f launches %d" % (d["value"], e["value"], e["sync_call_value"], e["padded_records_value"], e["padded_records_sync_value"], e["pcie_bound_qps"], e["pcie_bound_gbs"], e["pcie_frac"], d["gpu_launches"]))
PY
done
timeout 120 ./quadruped_control_b200/cpp/shim_latency 2>&1 | tail -1
