#!/bin/bash
# Round 2, GPU call 26: compute-stream count of the host pipeline, three runs each.
O=gpurun_out
mkdir -p $O
run() {
  timeout 300 python bench.py --no-secondary --steps 20 --warmup 3 2>/dev/null > $O/r2c26_tmp.json
  python - "$1" <<'PY'
import json, sys
d = json.load(open("gpurun_out/r2c26_tmp.json")); e = d["e2e"]
print("%-24s e2e(wire) %.3e sync %.3e | padded %.3e sync %.3e | frac %.3f" % (sys.argv[1], e["value"], e["sync_call_value"], e["padded_records_value"], e["padded_records_sync_value"], e["pcie_frac"]))
PY
}
for i in 1 2 3; do
for C in 2 3 4 6; do QPB_HOST_CSTREAMS=$C run "$C compute streams"; done
done | tee $O/r2c26_cstreams.txt
