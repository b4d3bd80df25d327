#!/bin/bash
# Round 2, GPU call 28: 1000-tick trajectory test, parameter fuzz over every kernel choice (cold, warm, forced paths).
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "thousand or warm" 2>&1 | tail -8 | cut -c1-250
timeout 1200 python tests/tools/fuzz_gpu.py 60 2048 77 2>&1 | tail -12 | cut -c1-300 | tee $O/r2c28_fuzz.txt
