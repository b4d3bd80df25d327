#!/bin/bash
# Round 2, GPU call 48: programmatic dependent launches of the loop / finishing passes against plain launches (alternating, one
# process), then every GPU test except the 32-s full-size config-5 one (verified in call 47; same kernels).
O=gpurun_out
mkdir -p $O
timeout 60 python tools/time_pdl.py 2> $O/r2c48_pdl.err | tee $O/r2c48_pdl_ab.txt; tail -2 $O/r2c48_pdl.err
timeout 85 python -m pytest tests -m gpu -q -p no:cacheprovider -k "not config5" 2>&1 | tail -6 | tee $O/r2c48_gpu_tests.log
