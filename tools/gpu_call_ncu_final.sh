#!/bin/bash
# Runs on the GPU box: launch list of the default bench command and one full capture of the balance kernel (final build).
O=gpurun_out
mkdir -p $O
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_final.csv \
    python bench.py --steps 10 --warmup 3 > $O/bench_under_ncu_final.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:balance_qp -s 3 -c 1 -f -o $O/prof_final \
    python tools/prof_run.py cfg2 6 > $O/prof_final.log 2>&1
python tools/ncu_digest.py $O/prof_final.ncu-rep 65536 > $O/prof_final_digest.txt 2>&1
rm -f $O/prof_final.ncu-rep
head -30 $O/prof_final_digest.txt | grep -E "duration|dram__bytes|issue_active|fp64|inst_executed.sum|registers"; grep -c balance_qp $O/launches_final.csv
