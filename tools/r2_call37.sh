#!/bin/bash
# Round 2, GPU call 37: set-up pass with its input staged through shared memory by coalesced cp.async (experiment build).
QPB_LIB=$PWD/scratch/libs/libqpb_stage2.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "profiles or mask or bad_input or three_entry or wire or warm or reference_sources" 2>&1 | tail -3
for L in quadruped_control_b200/libqpb200.so scratch/libs/libqpb_stage2.so quadruped_control_b200/libqpb200.so scratch/libs/libqpb_stage2.so; do
  a=$(QPB_LIB=$PWD/$L timeout 300 python bench.py --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); s=d['secondary']; print('cfg2 %.3e cfg3 %.3e warm %.3e' % (d['value'], s['cfg3']['value'], s['cfg2_warm_tick']['value']))")
  echo "$(basename $L): $a"
done | tee gpurun_out/r2c37_stage_inout.txt
