#!/bin/bash
# Round 2, GPU call 31: early finish for cold batches too? (A/B on config 2 / config 3)
for E in 0 1 0 1; do
  a=$(QPB_TPQ_EARLY_COLD=$E timeout 200 python bench.py --no-secondary --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3e err %.1e' % (d['value'], d['max_rel_grf_err_vs_oracle']))")
  c=$(QPB_TPQ_EARLY_COLD=$E timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3e err %.1e' % (d['value'], d['max_rel_grf_err_vs_oracle']))")
  echo "early_cold=$E  cfg2 $a  cfg3 $c"
done | tee gpurun_out/r2c31_early_cold.txt
