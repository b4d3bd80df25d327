#!/bin/bash
# Round 2, GPU call 38: finishing pass with its records staged through shared memory (against the build before it).
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_sanitizer_gpu.py -m gpu -q -x 2>&1 | tail -3
for L in scratch/libs/libqpb_prev.so quadruped_control_b200/libqpb200.so scratch/libs/libqpb_prev.so quadruped_control_b200/libqpb200.so; do
  a=$(QPB_LIB=$PWD/$L timeout 300 python bench.py --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); s=d['secondary']; print('cfg2 %.3e cfg3 %.3e warm %.3e' % (d['value'], s['cfg3']['value'], s['cfg2_warm_tick']['value']))")
  echo "$(basename $L): $a"
done | tee gpurun_out/r2c38_finish_staging.txt
timeout 300 python tools/time_warm.py 2>&1 | tail -3
