#!/bin/bash
# Round 2, GPU call 34: rolled epilogue with its arrays in shared memory (one-launch kernel, early-finish set-up).
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_sanitizer_gpu.py -m gpu -q -x 2>&1 | tail -4 | cut -c1-200 | sed "s/^/parity: /"
for i in 1 2; do
timeout 300 python bench.py --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); s=d['secondary']; print('cfg2 %.3e cfg3 %.3e warm %.3e e2e %.3e' % (d['value'], s['cfg3']['value'], s['cfg2_warm_tick']['value'], d['e2e']['value']))"
done
timeout 300 python tools/time_warm.py 2>&1 | tail -3
timeout 300 python tools/time_small_batches.py 2>&1 | tail -8
timeout 120 ./quadruped_control_b200/cpp/shim_latency 2>&1 | tail -1
