#!/bin/bash
# Round 2, GPU call 12: sanitizer tests after the G hand-off fix, small-batch latencies, default bench.
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_sanitizer_gpu.py -m gpu -q -x 2>&1 | tail -30 | cut -c1-220
timeout 300 python tools/time_small_batches.py 2>&1 | tail -8
timeout 200 python bench.py --no-secondary --steps 30 --warmup 5 2>/dev/null | cut -c1-120 | sed "s|^|cfg2: |"
timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 2>/dev/null | cut -c1-120 | sed "s|^|cfg3: |"
