"""Latency of small batches: CUDA-event time per launch of qpb_control_batch_packed for n = 1 ... 65536 records, for the
one-launch half-warp kernel and the range-space path with 1 / 2 / 4 lanes per QP (QPB_TPQ_MIN_N=0 forces it at every size).
Tells where the dispatch threshold (QPB_TPQ_MIN_N) belongs and what a loop iteration costs when the machine is empty."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from quadruped_control_b200 import OUT_DTYPE, default_params, lib, states

sizes = [1, 32, 148, 1024, 4096, 8192, 16384, 32768, 65536]
S = states.generate_states(max(sizes), 20260102)
d_in = torch.from_numpy(S.view(np.uint8).reshape(-1)).cuda()
d_out = torch.empty(len(S) * 256, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream()
print("mode        " + "".join(f"{n:>10d}" for n in sizes) + "   (us per launch)")
for name, env in (("half-warp", {"QPB_QPS_PER_WARP": "2"}), ("range lpq1", {"QPB_TPQ_MIN_N": "0", "QPB_TPQ_LPQ": "1"}),
                  ("range lpq2", {"QPB_TPQ_MIN_N": "0", "QPB_TPQ_LPQ": "2"}), ("range lpq4", {"QPB_TPQ_MIN_N": "0", "QPB_TPQ_LPQ": "4"})):
    for k in ("QPB_QPS_PER_WARP", "QPB_TPQ_MIN_N", "QPB_TPQ_LPQ"):
        os.environ.pop(k, None)
    os.environ.update(env)
    sol = lib.BalanceSolver(default_params(0.6))
    row = []
    for n in sizes:
        for _ in range(3):
            sol.control_packed(d_in, d_out, n, stream.cuda_stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record(stream)
        for _ in range(reps):
            sol.control_packed(d_in, d_out, n, stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        row.append(e0.elapsed_time(e1) * 1e3 / reps)
    it = d_out.cpu().numpy().view(OUT_DTYPE)["iters"]
    print(f"{name:12s}" + "".join(f"{t:10.1f}" for t in row) + f"   iters mean {it.mean():.1f} max {it.max()}")
    sol.close()
