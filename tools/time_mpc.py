"""Developer tool: time the MPC kernel on config 4 (device-resident records, CUDA events) and, with a
-DQPB_MPC_PROFILE build (QPB_LIB=...), print the share of cycles per phase.  Usage: python tools/time_mpc.py [n] [gaits]"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadruped_control_b200 import lib  # noqa: E402
from quadruped_control_b200.records import MPC_OUT_DTYPE  # noqa: E402
from quadruped_control_b200.states import generate_mpc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
gaits = sys.argv[2] if len(sys.argv) > 2 else "mixed"
R = generate_mpc(n, 20260104, gaits=gaits)
s = lib.MpcSolver(device=0)
d_in = torch.from_numpy(R.view(np.uint8).reshape(-1)).cuda()
d_out = torch.empty(n * MPC_OUT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
L = lib.load()
prof = hasattr(L, "qpb_mpc_debug_profile")
for _ in range(2):
    s.solve_packed(d_in, d_out, n, stream=st)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 24)()
if prof:
    L.qpb_mpc_debug_profile(buf)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    s.solve_packed(d_in, d_out, n, stream=st)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
out = d_out.cpu().numpy().view(MPC_OUT_DTYPE)
print(f"n={n} gaits={gaits}: {ms:.3f} ms per launch = {n / ms * 1e3:.4g} QP/s; status {np.bincount(out['status'])}; "
      f"iters mean {out['iters'].mean():.2f} max {out['iters'].max()}")
if prof:
    L.qpb_mpc_debug_profile(buf)
    v = np.array(list(buf), dtype=np.float64)
    cnt = max(v[5], 1.0)
    names = {16: "load", 17: "A", 18: "B", 19: "C", 20: "D", 0: "grad", 1: "sweep", 2: "start", 3: "loop-rest", 4: "io", 6: "select", 8: "nt", 9: "r", 10: "w", 11: "zt",
             12: "scalars", 13: "step", 14: "add", 15: "drop"}
    tot = sum(v[i] for i in names)
    print("cycles per QP: " + ", ".join(f"{nm} {v[i] / cnt:.0f} ({100 * v[i] / tot:.0f}%)" for i, nm in names.items()) + f"; total {tot / cnt:.0f}")
