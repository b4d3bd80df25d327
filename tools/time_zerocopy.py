"""Experiment: zero-copy host path (kernel reads/writes pinned host memory over PCIe) vs the staged pipeline."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from quadruped_control_b200 import lib, states, default_params, STATE_DTYPE, OUT_DTYPE
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
S = states.generate_states(n, 20260102)
pin_in = lib.PinnedBuffer(n, STATE_DTYPE); pin_out = lib.PinnedBuffer(n, OUT_DTYPE)
pin_in.array[:] = S
ref = None
for zc in ("0", "1"):
    os.environ["QPB_ZEROCOPY"] = zc
    sol = lib.BalanceSolver(default_params(0.6))
    for _ in range(3): sol.control_host(pin_in.array, pin_out.array)
    best = 1e9
    for rep in range(5):
        t0 = time.perf_counter()
        for _ in range(10): sol.control_host(pin_in.array, pin_out.array)
        best = min(best, (time.perf_counter() - t0) / 10)
    out = pin_out.array.copy()
    if ref is None: ref = out
    print(f"zero_copy={zc}: {best*1e3:.3f} ms  {n/best:.3e} QP/s  same bytes as staged: {out.tobytes()==ref.tobytes()}")
    sol.close()
