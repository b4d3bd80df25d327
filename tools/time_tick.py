"""Times qpb_tick_batch_packed (balance kernel + swing kernel) against qpb_control_batch_packed on config 3."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from quadruped_control_b200 import lib, states, default_params
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
S = states.generate_states(n, 20260103, masks="mixed"); SW = states.generate_swing(S, 5)
sol = lib.BalanceSolver(default_params(0.6))
d_s = torch.from_numpy(S.view(np.uint8).reshape(-1)).cuda(); d_w = torch.from_numpy(SW.view(np.uint8).reshape(-1)).cuda()
d_o = torch.empty(n * 256, dtype=torch.uint8, device="cuda")
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
tb = timeit(lambda: sol.control_packed(d_s, d_o, n)); tt = timeit(lambda: sol.tick_packed(d_s, d_w, d_o, n))
print(f"n={n}: balance only {tb:.3f} ms ({n/tb*1e3:.3e} robots/s); whole tick {tt:.3f} ms ({n/tt*1e3:.3e} ticks/s); swing kernel {tt-tb:.3f} ms "
      f"= {(n*(512+288+96))/(max(tt-tb,1e-9)*1e-3)/1e9:.0f} GB/s of 800 B read + 96 B written per robot")
