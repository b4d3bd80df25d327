#!/bin/bash
# Round 2 evidence: GPU tests, smoke, every bench line, ncu launch list + full digests of the three range-space kernels,
# small-batch latencies, shim latency.  Everything lands in gpurun_out/r2f_*; what is worth keeping is copied to profiles/.
O=gpurun_out
mkdir -p $O
( time timeout 1500 python -m pytest tests -m gpu -q ) 2>&1 | tail -6 > $O/r2f_gpu_tests.log; cat $O/r2f_gpu_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2f_smoke.log 2>&1; tail -2 $O/r2f_smoke.log
timeout 600 python bench.py > $O/r2f_bench_n1_cfg2.json 2> $O/r2f_bench.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2f_bench_reference_n1.json 2>> $O/r2f_bench.err
timeout 300 python bench.py --workload cfg3 --steps 10 --warmup 3 > $O/r2f_bench_n1_cfg3.json 2>> $O/r2f_bench.err
timeout 300 python bench.py --no-secondary --profile light > $O/r2f_bench_n1_cfg2_light.json 2>> $O/r2f_bench.err
timeout 300 python bench.py --no-secondary --profile stress > $O/r2f_bench_n1_cfg2_stress.json 2>> $O/r2f_bench.err
QPB_QPS_PER_WARP=2 timeout 300 python bench.py --no-secondary > $O/r2f_bench_n1_cfg2_halfwarp.json 2>> $O/r2f_bench.err
timeout 300 python bench.py --workload tick --steps 10 --warmup 3 > $O/r2f_bench_n1_tick.json 2>> $O/r2f_bench.err
for f in $O/r2f_bench_*.json; do echo "$(basename $f): $(cut -c1-110 $f)"; done; tail -2 $O/r2f_bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2f_launches_bench_steps10.csv python bench.py --no-secondary --steps 10 --warmup 3 > $O/r2f_bench_under_ncu.log 2>&1
grep -c "tpq_" $O/r2f_launches_bench_steps10.csv
for W in cfg2 cfg3; do for K in setup loop finish; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tpq_${K} -s 1 -c 1 -f -o $O/r2f_prof_${W}_$K python tools/prof_run.py $W 3 > $O/r2f_prof.log 2>&1
NQ=65536; [ $W = cfg3 ] && NQ=1048576
python tools/ncu_digest.py $O/r2f_prof_${W}_$K.ncu-rep $NQ > $O/r2f_ncu_full_${W}_${K}_digest.txt 2>&1
rm -f $O/r2f_prof_${W}_$K.ncu-rep
echo "== $W $K: $(grep -E 'gpu__time_duration' $O/r2f_ncu_full_${W}_${K}_digest.txt | awk '{print $2,$3}') issue $(grep issue_active $O/r2f_ncu_full_${W}_${K}_digest.txt | awk '{print $2}') dram r/w $(grep -E 'dram__bytes_(read|write)' $O/r2f_ncu_full_${W}_${K}_digest.txt | awk '{printf "%s %s ", $2, $3}')"
done; done
timeout 300 python tools/time_small_batches.py > $O/r2f_small_batches.txt 2>&1; cat $O/r2f_small_batches.txt | tail -6
timeout 120 ./quadruped_control_b200/cpp/shim_latency > $O/r2f_shim_latency.txt 2>&1; tail -6 $O/r2f_shim_latency.txt
