"""Hand-on of long QPs (QPB_TPQ_HAND = working-set changes after which the one-lane loop passes a QP to a second launch at
four lanes per QP; QPB_TPQ_HAND_TAIL = which build that launch is): device-resident microseconds per call on BASELINE
config 2 (65 536 all-stance records, eight batches in rotation: 403 MB > L2) and config 3 (1 048 576 mixed-contact
records), against the path without hand-on, with every output record compared bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from quadruped_control_b200 import OUT_DTYPE, default_params, lib, states

quick = "--quick" in sys.argv
if "--ncu" in sys.argv:  # under ncu: the launches of two config-2 calls without and with hand-on (cap = argv after --ncu)
    cap = sys.argv[sys.argv.index("--ncu") + 1]
    S = states.generate_states(65536, 20260102, masks="all4")
    d_in = torch.from_numpy(S.view(np.uint8).reshape(-1)).cuda()
    d_out = torch.empty(len(S) * 256, dtype=torch.uint8, device="cuda")
    for hand in ("0", cap):
        os.environ["QPB_TPQ_HAND"] = hand
        sol = lib.BalanceSolver(default_params(0.6))
        for _ in range(2):
            sol.control_packed(d_in, d_out, len(S))
        torch.cuda.synchronize()
        sol.close()
    sys.exit(0)
N2, NB = 65536, 8
S2 = states.generate_states(N2 * NB, 20260102, masks="all4")
S3 = states.generate_states(1048576, 20260103, masks="mixed")
d2 = torch.from_numpy(S2.view(np.uint8).reshape(-1)).cuda()
d3 = torch.from_numpy(S3.view(np.uint8).reshape(-1)).cuda()
o2 = torch.empty(N2 * NB * 256, dtype=torch.uint8, device="cuda")
o3 = torch.empty(len(S3) * 256, dtype=torch.uint8, device="cuda")
stream = torch.cuda.current_stream()
SB, OB = 512 * N2, 256 * N2


def run(sol):
    def cfg2():
        for b in range(NB):
            sol.control_packed(d2[b * SB:(b + 1) * SB], o2[b * OB:(b + 1) * OB], N2, stream.cuda_stream)

    def cfg3():
        sol.control_packed(d3, o3, len(S3), stream.cuda_stream)

    res = []
    for fn, calls, reps in ((cfg2, NB, 6), (cfg3, 1, 8)):
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
        res.append(best)
    return res, o2.cpu().numpy().view(OUT_DTYPE).copy(), o3.cpu().numpy().view(OUT_DTYPE).copy()


def diff(a, b):
    same = (a.view(np.uint8).reshape(len(a), -1) == b.view(np.uint8).reshape(len(b), -1)).all(axis=1)
    scale = np.maximum(np.abs(b["grf_body"]).max(axis=1), 1.0)
    return int((~same).sum()), float((np.abs(a["grf_body"] - b["grf_body"]).max(axis=1) / scale).max())


configs = [(0, 1)] + [(c, 1) for c in ((8, 12) if quick else (5, 7, 9, 11, 13, 16))] + [(c, 0) for c in ((8,) if quick else (7, 11))] + [(0, 1)]
base = None
print("hand tail   cfg2 us/call  QP/s      cfg3 us/call  QP/s      launches/call  records differing from no hand-on (cfg2, cfg3), max rel GRF diff")
for cap, tail in configs:
    os.environ["QPB_TPQ_HAND"] = str(cap)
    os.environ["QPB_TPQ_HAND_TAIL"] = str(tail)
    sol = lib.BalanceSolver(default_params(0.6))
    l0 = sol.launches
    sol.control_packed(d2[:SB], o2[:OB], N2, stream.cuda_stream)
    per_call = sol.launches - l0
    (t2, t3), r2, r3 = run(sol)
    sol.close()
    assert (r2["status"] == 0).all() and (r3["status"] == 0).all()
    if base is None:
        base = (r2, r3)
    (n2, e2), (n3, e3) = diff(r2, base[0]), diff(r3, base[1])
    print(f"{cap:4d} {tail:4d}   {t2:10.1f}  {N2 / t2 * 1e6:.3e}   {t3:10.1f}  {len(S3) / t3 * 1e6:.3e}   {per_call:6d}         "
          f"{n2} {n3}  {max(e2, e3):.1e}   iters max {r2['iters'].max()} {r3['iters'].max()}", flush=True)
