#!/bin/bash
# Round 2, GPU call 35: block round that also guesses the free x / y rows of heavily loaded QPs (QPB_START_SATURATE).
for L in quadruped_control_b200/libqpb200.so scratch/libs/libqpb_sat4.so scratch/libs/libqpb_sat6.so quadruped_control_b200/libqpb200.so scratch/libs/libqpb_sat6.so; do
  a=$(QPB_LIB=$PWD/$L timeout 300 python bench.py --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); s=d['secondary']; print('cfg2 %.3e (iters %.2f, err %.1e) cfg3 %.3e (iters %.2f, err %.1e) warm %.3e' % (d['value'], d['details']['iters_mean'], d['max_rel_grf_err_vs_oracle'], s['cfg3']['value'], s['cfg3']['iters_mean'], s['cfg3']['max_rel_grf_err_vs_oracle'], s['cfg2_warm_tick']['value']))")
  b=$(QPB_LIB=$PWD/$L timeout 300 python bench.py --no-secondary --profile light --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('light %.3e' % d['value'])")
  c=$(QPB_LIB=$PWD/$L timeout 300 python bench.py --no-secondary --profile stress --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('stress %.3e' % d['value'])")
  echo "$(basename $L): $a $b $c"
done | tee gpurun_out/r2c35_saturate.txt
