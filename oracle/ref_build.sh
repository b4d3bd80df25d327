#!/bin/bash
# Builds oracle/_ref/libqpb_ref.so: the REFERENCE's own hot-path sources, compiled where they lie under
# /root/reference, against the stand-in third-party headers of oracle/ref_stubs (see ref_glue.cpp).
# Test infrastructure only; outputs go to oracle/_ref/ (git-ignored, travels to the GPU box).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=/root/reference/quadruped_controller
[ -d "$REF" ] || { echo "reference not present: keeping the prebuilt oracle/_ref" >&2; exit 0; }
OUT="$HERE/_ref"
mkdir -p "$OUT"
SRC="$REF/src/quadruped_controller"
if [ "$OUT/libqpb_ref.so" -nt "$HERE/ref_glue.cpp" ] && [ "$OUT/libqpb_ref.so" -nt "$HERE/ref_glue_plan.cpp" ] && [ "$OUT/libqpb_ref.so" -nt "$HERE/plan_oracle.c" ] && \
   [ "$OUT/libqpb_ref.so" -nt "$HERE/ref_stubs/quadruped_controller/math/rigid3d.hpp" ] && [ "$OUT/libqpb_ref.so" -nt "$HERE/qpb_oracle.c" ] && \
   [ "$OUT/libqpb_ref.so" -nt "$HERE/ref_stubs/armadillo" ] && [ "$OUT/libqpb_ref.so" -nt "$HERE/ref_stubs/qpOASES.hpp" ]; then
  exit 0
fi
gcc -O2 -std=c99 -fPIC -ffp-contract=off -c "$HERE/qpb_oracle.c" -o "$OUT/qpb_oracle.o"
g++ -O2 -std=c++17 -fPIC -shared -ffp-contract=off -w \
    -I "$HERE/ref_stubs" -I "$REF/include" -I "$HERE" \
    "$SRC/balance_controller.cpp" "$SRC/kinematics.cpp" "$SRC/gait.cpp" "$SRC/math/numerics.cpp" "$SRC/joint_controller.cpp" \
    "$SRC/foot_planner.cpp" "$SRC/trajectory.cpp" "$HERE/ref_glue_plan.cpp" \
    "$HERE/ref_glue.cpp" "$OUT/qpb_oracle.o" -o "$OUT/libqpb_ref.so" -lm -lpthread
rm -f "$OUT/qpb_oracle.o"
echo "built $OUT/libqpb_ref.so"
