"""Per-CTA timeline of the loop and finishing passes of one device-resident call (developer build of the library with
-DQPB_TPQ_TIMELINE: every CTA stamps %globaltimer when it starts and when it ends; QPB_LIB points at that build).
Prints when the CTAs of the loop pass leave (the shape of its tail) and when the CTAs of the finishing pass start and end.
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC -DQPB_TPQ_TIMELINE -shared \
       -o scratch/libs/libqpb_timeline.so quadruped_control_b200/csrc/*.cu
  QPB_LIB=$PWD/scratch/libs/libqpb_timeline.so python tools/timeline.py [cfg2|cfg3]       (profiles/r02_overlap_ab.txt)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from quadruped_control_b200 import default_params, lib, states

which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
n, seed, masks = (65536, 20260102, "all4") if which == "cfg2" else (1048576, 20260103, "mixed")
S = states.generate_states(n, seed, masks=masks)
d_in = torch.from_numpy(S.view(np.uint8).reshape(-1)).cuda()
d_out = torch.empty(n * 256, dtype=torch.uint8, device="cuda")
L = lib.load()
buf = (ctypes.c_ulonglong * (4 * 8192))()
pct = (0, 10, 25, 50, 75, 90, 99, 100)
for _ in range(1):
    sol = lib.BalanceSolver(default_params(0.6))
    for _ in range(3):
        sol.control_packed(d_in, d_out, n)
    torch.cuda.synchronize()
    L.qpb_debug_timeline(buf)  # clears the stamps
    sol.control_packed(d_in, d_out, n)
    torch.cuda.synchronize()
    L.qpb_debug_timeline(buf)
    t = np.frombuffer(buf, dtype=np.uint64).reshape(4, 8192).astype(np.float64)
    ls, le, fs, fe = (t[k][t[k] > 0] for k in range(4))
    t0 = ls.min()
    us = lambda a: (a - t0) * 1e-3
    print(f"{which}: loop CTAs {len(ls)}, finish CTAs {len(fs)} (us after the first loop CTA started)")
    print("  percentile        " + "".join(f"{p:>8d}" for p in pct))
    for name, a in (("loop CTA start", ls), ("loop CTA end", le), ("finish CTA start", fs), ("finish CTA end", fe)):
        print(f"  {name:18s}" + "".join(f"{v:8.1f}" for v in np.percentile(us(a), pct)))
    print(f"  finish CTAs started before the last loop CTA ended: {(fs < le.max()).sum()} of {len(fs)};"
          f" ended before it: {(fe < le.max()).sum()};  loop pass {us(le).max():.1f} us, all done at {us(fe).max():.1f} us")
    sol.close()
