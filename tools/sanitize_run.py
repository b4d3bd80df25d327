"""Small run of every entry point for compute-sanitizer (memcheck / racecheck / initcheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from quadruped_control_b200 import lib, states, default_params, OUT_DTYPE
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1537
S = states.generate_states(n, 4, profile="stress", masks="mixed")
S["x"][3, 0] = np.nan
for mode in ("2", "1"):
    os.environ["QPB_QPS_PER_WARP"] = mode
    sol = lib.BalanceSolver(default_params(0.6))
    out = sol.control_host(S)
    d_in = torch.from_numpy(S.view(np.uint8).reshape(-1).copy()).cuda()
    d_out = torch.empty(n * 256, dtype=torch.uint8, device="cuda")
    sol.control_packed(d_in, d_out, n)
    dev = {k: torch.from_numpy(np.ascontiguousarray(S[k])).cuda() for k in ("Rwb", "Rwb_d", "x", "xdot", "w", "x_d", "xdot_d", "w_d", "feet", "contact", "q")}
    grf = torch.zeros(n, 12, dtype=torch.float64, device="cuda"); tau = torch.zeros_like(grf); st = torch.zeros(n, dtype=torch.int32, device="cuda")
    sol.control_split(n, dev["Rwb"], dev["Rwb_d"], dev["x"], dev["xdot"], dev["w"], dev["x_d"], dev["xdot_d"], dev["w_d"], dev["feet"], dev["contact"], dev["q"], grf, tau, st)
    torch.cuda.synchronize()
    assert d_out.cpu().numpy().view(OUT_DTYPE).tobytes() == out.tobytes()
    sol.fk_host(S["q"]); sol.jt_host(S["q"], out["grf_body"], S["contact"])
    sol.close()
print("sanitize_run ok", n)
