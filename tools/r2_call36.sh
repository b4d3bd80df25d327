#!/bin/bash
# Round 2, GPU call 36: synchronous host call with short first and last stages.
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "three_entry or wire or multi_device or async or small_batches or warm_batches" 2>&1 | tail -3
for R in 1 0 1 0; do
QPB_HOST_RAMP=$R timeout 300 python bench.py --no-secondary --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']
print('ramp=$R e2e(wire) %.3e sync %.3e | padded %.3e sync %.3e' % (e['value'], e['sync_call_value'], e['padded_records_value'], e['padded_records_sync_value']))"
done
