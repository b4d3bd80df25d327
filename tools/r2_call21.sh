#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 120 python tools/trace_host.py wire 2> $O/r2c21_trace_wire.txt; tail -36 $O/r2c21_trace_wire.txt
QPB_HOST_STAGES=4 timeout 120 python tools/trace_host.py wire 2> $O/r2c21_trace_wire_s4.txt; tail -20 $O/r2c21_trace_wire_s4.txt
