"""End-to-end sanity check without torch (starts in a second): BASELINE config 2 through qpb_control_batch_wire_host_async from
pinned wire records, two batches in flight, wall clock around K calls + qpb_host_sync.  Prints QP/s."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from quadruped_control_b200 import WIRE_OUT_DTYPE, WIRE_STATE_DTYPE, default_params, lib, states, to_wire

n, K = 65536, 40
S = to_wire(states.generate_states(n, 20260102, masks="all4"))
sol = lib.BalanceSolver(default_params(0.6))
pin_i = [lib.PinnedBuffer(n, WIRE_STATE_DTYPE) for _ in range(2)]
pin_o = [lib.PinnedBuffer(n, WIRE_OUT_DTYPE) for _ in range(2)]
for b in pin_i:
    b.array[:] = S
for rep in range(3):
    if rep:
        t0 = time.perf_counter()
    for k in range(K if rep else 6):
        sol.control_wire_host_async(pin_i[k % 2].array, pin_o[k % 2].array)
    sol.host_sync()
    if rep:
        dt = time.perf_counter() - t0
        print(f"e2e async wire records: {n * K / dt:.4e} QP/s ({dt / K * 1e3:.3f} ms per batch), status ok {(pin_o[0].array['status'] == 0).all()}", flush=True)
sol.close()
