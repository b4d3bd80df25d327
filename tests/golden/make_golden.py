"""Regenerates balance_golden.npz from the CPU oracle (run from the repo root)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from quadruped_control_b200 import default_params, states  # noqa: E402


def golden_states():
    parts = [states.stance_state()]
    for prof in ("default", "light", "stress"):
        parts.append(states.generate_states(16, 7001, profile=prof, masks="all4"))
        parts.append(states.generate_states(16, 7002, profile=prof, masks="mixed"))
    every = states.generate_states(5 * 16, 7003)
    for i in range(len(every)):  # all 16 contact masks, including 0- and 1-foot
        m = i % 16
        every["contact"][i] = [(m >> k) & 1 for k in range(4)]
    parts.append(every)
    return np.concatenate(parts)


if __name__ == "__main__":
    S = golden_states()
    out = {}
    for mu in (0.6, 0.8):
        p = default_params(mu)
        O = oracle.control_batch(p, S, 1)
        out[f"grf_mu{mu}"] = O["grf_body"]
        out[f"tau_mu{mu}"] = O["tau"]
        out[f"status_mu{mu}"] = O["status"]
    np.savez_compressed(os.path.join(os.path.dirname(__file__), "balance_golden.npz"),
                        states=S.view(np.uint8).reshape(len(S), -1), **out)
    print("wrote", len(S), "states")
