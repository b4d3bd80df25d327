"""Do the kernel chains of two batches overlap when they are launched on different streams?  Device-resident batches of m
records, k streams, round-robin: microseconds per batch against the one-stream figure."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from quadruped_control_b200 import default_params, lib, states
m = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
os.environ["QPB_TPQ_MIN_N"] = "0"
sol = lib.BalanceSolver(default_params(0.6))
S = states.generate_states(m * 8, 20260102)
d_in = [torch.from_numpy(S[i * m:(i + 1) * m].view(np.uint8).reshape(-1).copy()).cuda() for i in range(8)]
d_out = [torch.empty(m * 256, dtype=torch.uint8, device="cuda") for _ in range(8)]
for k in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(k)]
    def go(reps):
        for r in range(reps):
            i = r % 8
            sol.control_packed(d_in[i], d_out[i], m, streams[r % k].cuda_stream)
    go(16)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams:
        s.wait_stream(torch.cuda.current_stream())
    go(64)
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    print(f"{m} records per batch, {k} stream(s): {e0.elapsed_time(e1) * 1e3 / 64:7.1f} us per batch")
