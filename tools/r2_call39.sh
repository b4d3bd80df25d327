#!/bin/bash
# Round 2, GPU call 39: hand-on of long QPs measured (caps x builds of the second loop launch), every GPU test, launch list
# of a config-2 call without / with hand-on, default bench line.
O=gpurun_out
mkdir -p $O
timeout 300 python tools/time_hand.py > $O/r2c39_hand_ab.txt 2> $O/r2c39_hand_ab.err
tail -14 $O/r2c39_hand_ab.txt
timeout 560 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=6 2>&1 | tail -16 > $O/r2c39_gpu_tests.log
cat $O/r2c39_gpu_tests.log
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $O/r2c39_launches_hand.csv python tools/time_hand.py --ncu 9 > /dev/null 2>&1
grep -c tpq $O/r2c39_launches_hand.csv
timeout 200 python bench.py --steps 50 --warmup 5 > $O/r2c39_bench.json 2> $O/r2c39_bench.err
cut -c1-300 $O/r2c39_bench.json; tail -2 $O/r2c39_bench.err
