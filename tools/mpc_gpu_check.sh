timeout 600 python -m pytest tests/test_mpc.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/mpc_tests.log
timeout 200 python tools/time_mpc.py 65536 mixed > gpurun_out/mpc_time.log 2>&1
export QPB_LIB=$PWD/quadruped_control_b200/libqpb200_prof.so
timeout 200 python tools/time_mpc.py 16384 mixed >> gpurun_out/mpc_time.log 2>&1
for g in stand trot crawl; do timeout 200 python tools/time_mpc.py 8192 $g >> gpurun_out/mpc_time.log 2>&1; done
cat gpurun_out/mpc_tests.log gpurun_out/mpc_time.log
