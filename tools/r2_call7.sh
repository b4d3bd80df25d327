#!/bin/bash
# Round 2, GPU call 7: ncu full captures of the three range-space kernels (cfg3, lpq1) + e2e sanity of the new host pipeline.
O=gpurun_out
mkdir -p $O
for K in loop setup finish; do
QPB_TPQ_LPQ=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:tpq_${K} -s 1 -c 1 -f -o $O/r2c7_prof_cfg3_$K python tools/prof_run.py cfg3 3 > $O/r2c7_prof.log 2>&1
python tools/ncu_digest.py $O/r2c7_prof_cfg3_$K.ncu-rep 1048576 > $O/r2c7_prof_cfg3_${K}_digest.txt 2>&1
echo "== $K"; grep -E "duration|issue_active|fp64|warps_active|stalled|thread_inst|dram__bytes|local" $O/r2c7_prof_cfg3_${K}_digest.txt
rm -f $O/r2c7_prof_cfg3_$K.ncu-rep
done
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
timeout 200 python bench.py --steps 30 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg2 value', d['value'], 'e2e', d['e2e'])"
