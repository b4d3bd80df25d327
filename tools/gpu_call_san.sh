#!/bin/bash
O=gpurun_out
mkdir -p $O
for t in memcheck racecheck initcheck; do
  timeout 250 compute-sanitizer --tool $t --print-limit 8 python tools/sanitize_run3.py > $O/san3_$t.log 2>&1
  grep -E "SUMMARY|sanitize_run3 ok|Error|error" $O/san3_$t.log | head -5
done
