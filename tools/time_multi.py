"""Developer tool: qpb_multi_control_batch_host (single process, one shard per device) against the single-handle call.
Pinned host buffers, 65 536 config-2 states per device (weak scaling), wall clock around the synchronous call."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from quadruped_control_b200 import OUT_DTYPE, STATE_DTYPE, default_params, lib, states  # noqa: E402

per_dev = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = 30
params = default_params(0.6)
ndev = torch.cuda.device_count()
single = lib.BalanceSolver(params, device=0)
for g in sorted({1, 2, ndev} & set(range(1, ndev + 1))):
    n = per_dev * g
    S = states.generate_states(n, 20260102)
    pin_in, pin_out = lib.PinnedBuffer(n, STATE_DTYPE), lib.PinnedBuffer(n, OUT_DTYPE)
    pin_in.array[:] = S
    multi = lib.MultiBalanceSolver(params, devices=list(range(g)))
    for _ in range(5):
        multi.control_host(pin_in.array, pin_out.array)
    t0 = time.perf_counter()
    for _ in range(reps):
        multi.control_host(pin_in.array, pin_out.array)
    dt = (time.perf_counter() - t0) / reps
    ref = single.control_host(S[: min(n, 20000)])
    same = pin_out.array[: len(ref)].tobytes() == ref.tobytes()
    print(f"qpb_multi_control_batch_host, {g} device(s), {n} states: {dt * 1e3:.3f} ms per call = {n / dt:.3e} QP/s"
          f" end to end (pinned host buffers); first {len(ref)} results identical to the single-handle call: {same}")
    if g == 1:
        for _ in range(5):
            single.control_host(pin_in.array, pin_out.array)
        t0 = time.perf_counter()
        for _ in range(reps):
            single.control_host(pin_in.array, pin_out.array)
        dt1 = (time.perf_counter() - t0) / reps
        print(f"qpb_control_batch_host (single handle), {n} states: {dt1 * 1e3:.3f} ms per call = {n / dt1:.3e} QP/s")
    multi.close()
    pin_in.free()
    pin_out.free()
