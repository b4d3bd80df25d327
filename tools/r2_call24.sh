#!/bin/bash
# Round 2, GPU call 24: host pipeline with one upload and one download stream: full suite, end to end, timeline.
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | cut -c1-200 | sed "s/^/suite: /"
for S in 8 12; do
QPB_HOST_STAGES=$S timeout 300 python bench.py --no-secondary --steps 20 --warmup 3 2>/dev/null > $O/r2c24_bench_$S.json
python - <<PY
import json
d = json.load(open("gpurun_out/r2c24_bench_$S.json")); e = d["e2e"]
print("stages=$S value %.3e e2e(wire) %.3e sync %.3e | padded %.3e sync %.3e | bound %.3e = %.1f GB/s frac %.3f launches %d" % (d["value"], e["value"], e["sync_call_value"], e["padded_records_value"], e["padded_records_sync_value"], e["pcie_bound_qps"], e["pcie_bound_gbs"], e["pcie_frac"], d["gpu_launches"]))
PY
done
QPB_HOST_TRACE=1 timeout 120 python tools/trace_host.py wire 2> $O/r2c24_trace_wire.txt | tail -2; tail -50 $O/r2c24_trace_wire.txt | head -14
timeout 200 python bench.py --workload tick --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('tick value %.3e e2e %.3e' % (d['value'], d['e2e']['value']))"
