#!/bin/bash
# Round 2, GPU call 33: finish<4> in the finishing pass and the early-finish set-up: parity, bench, warm sizes.
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4 | cut -c1-200 | sed "s/^/parity: /"
for i in 1 2; do
timeout 300 python bench.py --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); s=d['secondary']; print('cfg2 %.3e cfg3 %.3e warm %.3e e2e %.3e' % (d['value'], s['cfg3']['value'], s['cfg2_warm_tick']['value'], d['e2e']['value']))"
done
timeout 300 python tools/time_warm.py 2>&1 | tail -3
