#!/bin/bash
# Round 2, GPU call 32: finish() with its leg loop unrolled (no local-memory arrays) against the rolled loop.
for L in quadruped_control_b200/libqpb200.so scratch/libs/libqpb_fu4.so quadruped_control_b200/libqpb200.so scratch/libs/libqpb_fu4.so; do
  a=$(QPB_LIB=$PWD/$L timeout 300 python bench.py --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); s=d['secondary']; print('cfg2 %.3e cfg3 %.3e warm %.3e' % (d['value'], s['cfg3']['value'], s['cfg2_warm_tick']['value']))")
  echo "$(basename $L): $a"
done | tee gpurun_out/r2c32_finish_unroll.txt
for L in quadruped_control_b200/libqpb200.so scratch/libs/libqpb_fu4.so; do QPB_LIB=$PWD/$L timeout 200 python tools/time_small_batches.py 2>&1 | grep "one warm\|one-launch" | sed "s/^/$(basename $L) /"; done
