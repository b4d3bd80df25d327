"""Timeline of the host pipeline: a few asynchronous 65 536-record batches, wire or padded records.  With QPB_HOST_TRACE=1
in the environment the library prints per-stage event times; this script prints how long each asynchronous call kept the
CPU (a call that only queues work returns in a few hundred microseconds)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from quadruped_control_b200 import OUT_DTYPE, STATE_DTYPE, WIRE_OUT_DTYPE, WIRE_STATE_DTYPE, default_params, lib, states, to_wire
wire = len(sys.argv) > 1 and sys.argv[1] == "wire"
n = 65536
S = states.generate_states(n, 20260102)
sol = lib.BalanceSolver(default_params(0.6))
pi = [lib.PinnedBuffer(n, WIRE_STATE_DTYPE if wire else STATE_DTYPE) for _ in range(2)]
po = [lib.PinnedBuffer(n, WIRE_OUT_DTYPE if wire else OUT_DTYPE) for _ in range(2)]
for b in pi:
    if wire:
        to_wire(S, b.array)
    else:
        b.array[:] = S
call = sol.control_wire_host_async if wire else sol.control_host_async
for rep in range(3):
    t = [time.perf_counter()]
    for i in range(6):
        call(pi[i % 2].array, po[i % 2].array)
        t.append(time.perf_counter())
    sol.host_sync()
    t.append(time.perf_counter())
    print("rep", rep, "wire" if wire else "padded", "async calls took (us):", " ".join(f"{(b - a) * 1e6:.0f}" for a, b in zip(t[:-2], t[1:-1])),
          "| sync", f"{(t[-1] - t[-2]) * 1e6:.0f}", "| total", f"{(t[-1] - t[0]) * 1e6:.0f}")
