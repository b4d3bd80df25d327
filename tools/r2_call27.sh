#!/bin/bash
# Round 2, GPU call 27: large device-resident batches cut into parts on helper streams (fork / join): parity incl. CUDA-graph
# replay, A/B on config 2 / config 3.
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -12 | cut -c1-250 | sed "s/^/parity: /"
for P in 1 2 3 4 6 8; do
  a=$(QPB_SPLIT=$P timeout 200 python bench.py --no-secondary --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3e' % d['value'])")
  b=$(QPB_SPLIT=$P timeout 200 python bench.py --no-secondary --steps 30 --warmup 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3e' % d['value'])")
  c=$(QPB_SPLIT=$P timeout 200 python bench.py --workload cfg3 --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3e' % d['value'])")
  echo "parts=$P  cfg2 $a $b  cfg3 $c"
done | tee $O/r2c27_split_ab.txt
