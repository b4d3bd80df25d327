#!/bin/bash
# Round 2, GPU call 25: do the kernel chains of consecutive stages overlap?  Compute-stream count and hardware connections.
O=gpurun_out
mkdir -p $O
run() {
  timeout 300 python bench.py --no-secondary --steps 20 --warmup 3 2>/dev/null > $O/r2c25_tmp.json
  python - "$1" <<'PY'
import json, sys
d = json.load(open("gpurun_out/r2c25_tmp.json")); e = d["e2e"]
print("%-40s e2e(wire) %.3e sync %.3e | padded %.3e sync %.3e | frac %.3f" % (sys.argv[1], e["value"], e["sync_call_value"], e["padded_records_value"], e["padded_records_sync_value"], e["pcie_frac"]))
PY
}
run "default (6 compute streams, 8 conn)"
CUDA_DEVICE_MAX_CONNECTIONS=32 run "32 connections"
QPB_HOST_CSTREAMS=3 run "3 compute streams"
QPB_HOST_CSTREAMS=2 run "2 compute streams"
CUDA_DEVICE_MAX_CONNECTIONS=32 QPB_HOST_STAGES=12 run "32 connections, 12 stages"
CUDA_DEVICE_MAX_CONNECTIONS=32 QPB_HOST_STAGES=16 run "32 connections, 16 stages"
