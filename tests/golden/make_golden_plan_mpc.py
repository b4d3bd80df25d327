"""Regenerates plan_golden.npz (from the REFERENCE's own foot_planner.cpp / trajectory.cpp compiled in oracle/_ref; needs
/root/reference or a prebuilt oracle/_ref) and mpc_golden.npz (from the CPU oracle: the reference has no MPC code).
Run from the repo root:  python tests/golden/make_golden_plan_mpc.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from quadruped_control_b200 import states  # noqa: E402
from quadruped_control_b200.records import PLAN_DTYPE, SWING_DTYPE, default_mpc_params, default_plan_params  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def plan_inputs():
    S = np.concatenate([states.generate_states(96, 8101, masks="mixed"), states.generate_states(32, 8102, masks="all4")])
    S["contact"][96::2] = 0  # flight phases
    rng = np.random.default_rng(8103)
    plan = np.zeros(len(S), dtype=PLAN_DTYPE)
    plan["phase"] = rng.uniform(0.6, 1.05, size=(len(S), 4))
    plan["replan"] = rng.integers(0, 2, size=(len(S), 4))
    plan["p_start"] = rng.normal(0, 0.3, size=(len(S), 12))
    plan["p_final"] = rng.normal(0, 0.3, size=(len(S), 12))
    return S, plan


if __name__ == "__main__":
    assert oracle.ref_available(), "oracle/_ref is needed for the planner fixture"
    S, plan = plan_inputs()
    plan_out, sw = plan.copy(), np.zeros(len(S), dtype=SWING_DTYPE)
    oracle.ref_plan_batch(default_plan_params(), S, plan_out, sw)
    np.savez_compressed(os.path.join(HERE, "plan_golden.npz"), states=S.view(np.uint8).reshape(len(S), -1),
                        plan_in=plan.view(np.uint8).reshape(len(S), -1), plan_out=plan_out.view(np.uint8).reshape(len(S), -1),
                        foot_ref_pos=sw["foot_ref_pos"], foot_ref_vel=sw["foot_ref_vel"])
    R = np.concatenate([states.generate_mpc(36, 8201), states.generate_mpc(12, 8202, scale=3.0)])
    R["contact"][0] = 0
    out = oracle.mpc_batch(default_mpc_params(), R, os.cpu_count() or 1)
    np.savez_compressed(os.path.join(HERE, "mpc_golden.npz"), recs=R.view(np.uint8).reshape(len(R), -1), U=out["U"], status=out["status"])
    print("wrote", len(S), "planner cases and", len(R), "MPC cases")
