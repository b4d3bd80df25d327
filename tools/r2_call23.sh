#!/bin/bash
O=gpurun_out
mkdir -p $O
echo "== no trace"; timeout 120 python tools/trace_host.py wire 2>&1 | tail -3
echo "== no trace, padded"; timeout 120 python tools/trace_host.py 2>&1 | tail -3
echo "== no trace, 32 connections"; CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 120 python tools/trace_host.py wire 2>&1 | tail -3
echo "== trace, 32 connections"; CUDA_DEVICE_MAX_CONNECTIONS=32 QPB_HOST_TRACE=1 timeout 120 python tools/trace_host.py wire 2>&1 | tail -52 | head -12
