"""Times qpb_control_batch_host (pinned host buffers) for several pipeline chunk sizes."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from quadruped_control_b200 import lib, states, default_params, STATE_DTYPE, OUT_DTYPE
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
S = states.generate_states(n, 20260102)
pin_in = lib.PinnedBuffer(n, STATE_DTYPE); pin_out = lib.PinnedBuffer(n, OUT_DTYPE)
pin_in.array[:] = S
for chunk in (2048, 4096, 8192, 16384):
    os.environ["QPB_HOST_CHUNK"] = str(chunk)
    sol = lib.BalanceSolver(default_params(0.6))
    for _ in range(3): sol.control_host(pin_in.array, pin_out.array)
    best = 1e9
    for rep in range(5):
        t0 = time.perf_counter()
        for _ in range(10): sol.control_host(pin_in.array, pin_out.array)
        best = min(best, (time.perf_counter() - t0) / 10)
    print(f"chunk {chunk:6d}: {best*1e3:.3f} ms  {n/best:.3e} QP/s  H2D-equivalent {n*512/best/1e9:.1f} GB/s  ok={bool((pin_out.array['status']==0).all())}")
    sol.close()
# raw H2D / D2H bandwidth of this box for reference
import torch
h = torch.empty(n * 512, dtype=torch.uint8).pin_memory(); d = torch.empty(n * 512, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    print(f"raw {name}: {n*512/dt/1e9:.1f} GB/s for {n*512/1e6:.1f} MB")
# bidirectional test: H2D and D2H on two streams at once
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n * 256, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n * 256, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
print(f"bidirectional: {dt*1e3:.3f} ms per (33.6 MB up + 16.8 MB down) -> {n*768/dt/1e9:.1f} GB/s combined")
print("asyncEngineCount", torch.cuda.get_device_properties(0))
