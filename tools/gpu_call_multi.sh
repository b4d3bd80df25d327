#!/bin/bash
# Runs on a multi-GPU box: the single-process multi-device calls on real devices, then the torchrun bench line.
O=gpurun_out
mkdir -p $O
nvidia-smi -L > $O/multi_gpus.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "multi_device" 2>&1 | tail -5 > $O/multi_tests.log
timeout 200 python tools/time_multi.py > $O/multi_time.txt 2>&1
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > $O/bench_n${N}_multi.json 2> $O/bench_multi.err
cat $O/multi_gpus.txt $O/multi_tests.log $O/multi_time.txt; cut -c1-400 $O/bench_n${N}_multi.json; tail -3 $O/bench_multi.err
