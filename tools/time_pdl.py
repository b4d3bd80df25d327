"""Programmatic dependent launches of the loop and finishing passes (QPB_TPQ_PDL, default on) against plain launches, in one
process, alternating; each line: device-resident timings for the build QPB_LIB points at (default: the in-tree one):
BASELINE config 2 (65 536 all-stance records, eight batches in rotation: 403 MB > L2), config 3 (1 048 576 mixed-contact
records) and the same records one tick later with their working sets (warm), microseconds per call; plus a checksum of
the results so that two builds can be compared.  For A/B runs of compile-time variants (scratch/libs/*.so)."""
import hashlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from quadruped_control_b200 import OUT_DTYPE, default_params, lib, states

N2, NB = 65536, 8
S2 = states.generate_states(N2 * NB, 20260102, masks="all4")
S3 = states.generate_states(1048576, 20260103, masks="mixed")
d2 = torch.from_numpy(S2.view(np.uint8).reshape(-1)).cuda()
d3 = torch.from_numpy(S3.view(np.uint8).reshape(-1)).cuda()
o2 = torch.empty(N2 * NB * 256, dtype=torch.uint8, device="cuda")
o3 = torch.empty(len(S3) * 256, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream()
SB, OB = 512 * N2, 256 * N2


def timed(fn, calls, reps):
    for _ in range(2):
        fn()
    best = 1e30
    for _ in range(3):  # best of three timed regions
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / (reps * calls))
    return best


for pdl in ("0", "1", "0", "1"):
    os.environ["QPB_TPQ_PDL"] = pdl
    sol = lib.BalanceSolver(default_params(0.6))


    def cfg2():
        for b in range(NB):
            sol.control_packed(d2[b * SB:(b + 1) * SB], o2[b * OB:(b + 1) * OB], N2, stream.cuda_stream)


    t2 = timed(cfg2, NB, 6)
    t3 = timed(lambda: sol.control_packed(d3, o3, len(S3), stream.cuda_stream), 1, 8)
    r2, r3 = o2.cpu().numpy().view(OUT_DTYPE).copy(), o3.cpu().numpy().view(OUT_DTYPE).copy()
    small = [timed(lambda m=m: sol.control_packed(d3, o3, m, stream.cuda_stream), 1, 20) for m in (16384, 131072)]
    ok = bool((r2["status"] == 0).all() and (r3["status"] == 0).all())
    digest = hashlib.sha1(r2.tobytes() + r3.tobytes()).hexdigest()[:12]
    name = "QPB_TPQ_PDL=" + pdl
    print(f"{name:34s} cfg2 {t2:7.1f} us {N2 / t2 * 1e6:.3e} QP/s   cfg3 {t3:7.1f} us {len(S3) / t3 * 1e6:.3e} QP/s   16384 / 131072 mixed: {small[0]:.1f} / {small[1]:.1f} us   ok={ok} results sha1 {digest}", flush=True)
    sol.close()
