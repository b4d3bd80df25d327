#!/bin/bash
# Round 2, GPU call 8: the default bench line (with the secondary block, asynchronous e2e, PCIe bound) and the reference arm.
O=gpurun_out
mkdir -p $O
timeout 600 python bench.py > $O/r2c8_bench.json 2> $O/r2c8_bench.err; echo "rc=$?"; tail -3 $O/r2c8_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c8_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'launches',d['gpu_launches'])
print('roofline',{k:d['roofline'][k] for k in ('achieved','peak','frac','kernel')}, d['roofline']['fp64'])
print('e2e',{k:v for k,v in d['e2e'].items() if k not in ('api','sync_api','pcie_note')})
print('secondary',json.dumps(d['secondary'])[:1500])
print('cpu',d['cpu_baseline']['value'],d['max_rel_grf_err_vs_oracle'], d['config'])
PY
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | cut -c1-700
