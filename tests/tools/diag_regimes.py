import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, oracle
from quadruped_control_b200 import default_params, states, lib
rng = np.random.default_rng(12)
S = states.generate_states(3000, 91, profile="stress", masks="mixed")
S["x_d"][:, 2] = S["x"][:, 2] - rng.uniform(0.0, 0.6, len(S))
S2 = states.generate_states(3000, 92, profile="default", masks="mixed")
def mk(mu, **kw):
    p = default_params(mu)
    for k, v in kw.items(): setattr(p, k, v)
    return p
pw = default_params(0.6); pw.W[:] = (1e-1 * np.eye(12)).ravel().tolist()
for name, p, Sin in (("apex", mk(0.6, fzmin=0.0), S), ("fzfixed", mk(0.6, fzmin=35.0, fzmax=35.0), S), ("mutiny", mk(0.01), S),
                     ("mularge", mk(5.0), S), ("heavy", mk(0.6, mass=60.0), S2), ("Wlarge", pw, S2)):
    ref = oracle.control_batch(p, Sin, 8)
    for mode in ("2", "1"):
        os.environ["QPB_QPS_PER_WARP"] = mode
        sol = lib.BalanceSolver(p); out = sol.control_host(Sin); sol.close()
        bad = np.nonzero(out["status"] != ref["status"])[0]
        ok = out["status"] == 0
        err = (np.abs(out["grf_body"][ok] - ref["grf_body"][ok]).max(axis=1) / np.maximum(np.abs(ref["grf_body"][ok]).max(axis=1), 1)).max()
        print(name, "mode", mode, "status", np.bincount(out["status"], minlength=3), "mismatch", len(bad), "iters max", out["iters"].max(), "err(ok)", f"{err:.2e}")
        for b in bad[:3]:
            print("   idx", b, "status", out["status"][b], "iters", out["iters"][b], "ref iters", ref["iters"][b], "contact", Sin["contact"][b], "ref f", np.round(ref["grf_body"][b], 3))
