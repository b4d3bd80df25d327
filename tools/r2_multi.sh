#!/bin/bash
# Round 2, multi-GPU box: bench.py under torchrun at every N the box has (device-resident, asynchronous end-to-end, PCIe bound,
# secondary configs), plus the single-process multi-device test.
O=gpurun_out
mkdir -p $O
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"; nvidia-smi topo -m 2>/dev/null | head -14 > $O/r2m_topo.txt; lscpu | grep -E "NUMA|Socket|^CPU\(s\)" >> $O/r2m_topo.txt
for N in ${NLIST:-1 2 4 8}; do
  [ $N -gt $NG ] && continue
  if [ $N -eq 1 ]; then
    timeout 600 python bench.py --steps 20 --warmup 3 > $O/r2m_bench_n1.json 2> $O/r2m_bench_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) bench.py --gpus $N --steps 20 --warmup 3 > $O/r2m_bench_n$N.json 2> $O/r2m_bench_n$N.err
  fi
  python - $O/r2m_bench_n$N.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith('{')][-1]); e=d['e2e']
    print(f"N={d['n_gpus']}: value {d['value']:.3e}  e2e {e['value']:.3e} (sync {e['sync_call_value']:.3e})  pcie bound {e['pcie_bound_qps']:.3e} = {e['pcie_bound_gbs']:.0f} GB/s  frac {e['pcie_frac']:.2f}  secondary "+", ".join(f"{k} {v['value']:.3e}" for k,v in (d.get('secondary') or {}).items() if 'value' in v), (d.get('secondary') or {}).get('allgather_outputs'))
except Exception as ex:
    print(sys.argv[1], 'ERR', ex)
PY
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "multi_device" 2>&1 | tail -2
