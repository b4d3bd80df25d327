#!/bin/bash
# Static kernel evidence (no GPU needed): ptxas resource usage and SASS opcode histograms of the balance kernels of the
# built library -> profiles/r02_sass_balance.txt.  Run after __graft_entry__.build().
set -e
cd "$(dirname "$0")/.."
LIB=quadruped_control_b200/libqpb200.so
OUT=profiles/r02_sass_balance.txt
{
echo "# ptxas resource usage (cuobjdump -res-usage $LIB), balance kernels"
cuobjdump -res-usage $LIB 2>/dev/null | grep -A1 -E "tpq_|balance_qp|wire_" | grep -E "Function|REG" | sed 's/ Function \(.*\):/\1/' | paste - - | sed 's/^ *//' | cut -c1-260
echo
for K in tpq_setup_kernelINS_8PackedIOELb0 tpq_setup_kernelINS_8PackedIOELb1 tpq_one_kernelINS_8PackedIOELb1 tpq_one_kernelINS_8PackedIOELb0 wire_unpack_kernel wire_pack_kernel tpq_loop_kernelILi1 tpq_loop_kernelILi2 tpq_loop_kernelILi4 tpq_finish_kernelINS_8PackedIOELb0 balance_qp_kernel16INS_8PackedIO; do
  F=$(cuobjdump -sass $LIB 2>/dev/null | grep -o "Function : [_A-Za-z0-9]*${K}[_A-Za-z0-9]*" | head -1 | sed 's/Function : //')
  [ -z "$F" ] && continue
  echo "# SASS opcode histogram: $F"
  cuobjdump -sass -fun "$F" $LIB 2>/dev/null | grep -oE "^\s*/\*[0-9a-f]{4,}\*/\s+(@!?U?P[0-9T] )?[A-Z0-9_.]+" | awk '{print $NF}' > /tmp/_ops.txt
  echo "total instructions: $(wc -l < /tmp/_ops.txt)"
  sed 's/\..*//' /tmp/_ops.txt | sort | uniq -c | sort -rn | head -28 | awk '{printf "%s %s, ", $2, $1} END {print ""}'
  echo "memory ops by width: $(grep -E '^(LDG|STG|LDS|STS|LDL|STL|LDGSTS)' /tmp/_ops.txt | sort | uniq -c | sort -rn | awk '{printf "%s %s, ", $2, $1}')"
  echo "tensor-core / wgmma opcodes (none expected: no dense contraction): $(grep -cE 'MMA|UTC' /tmp/_ops.txt || true)"
  echo
done
} > $OUT
wc -l $OUT; head -12 $OUT | cut -c1-200
