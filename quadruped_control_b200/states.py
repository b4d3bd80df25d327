"""Synthetic batched robot states, index-addressable.

Record ``i`` of a stream depends only on ``(seed, i)``: every record consumes exactly
``DRAWS_PER_RECORD`` 64-bit Philox outputs, so a rank that owns ``[lo, hi)`` advances the
counter to ``lo`` and generates its shard without touching the rest.  The same function feeds
the CUDA path, the parity tests and the CPU baseline, so all three see identical bytes.

Foot positions come from the reference's own forward kinematics (kinematics.cpp:81-103), which is
how the caller produces ``foot_map`` (commander_node.cpp:383-384).
"""
import numpy as np

from .records import STATE_DTYPE, Params, default_params

DRAWS_PER_RECORD = 64  # 16 Philox4x64 blocks

# stance pose, quadruped_controller/config/gait_visualizer.yaml:47-50 (RL FL RR FR)
STANCE_Q = np.array([0.056, 0.90, -1.94, 0.056, 0.90, -1.94, -0.056, 0.90, -1.94, -0.056, 0.90, -1.94])

PROFILES = {
    # attitude error bound (rad), position box (m), velocity sigmas, angular-velocity sigmas
    "default": dict(att=0.05, pos=0.05, vel=0.3, vel_d=0.3, w=0.5, w_d=0.2),
    "light": dict(att=0.002, pos=0.01, vel=0.02, vel_d=0.02, w=0.01, w_d=0.01),
    "stress": dict(att=0.3, pos=0.05, vel=0.3, vel_d=0.3, w=0.5, w_d=0.2),
}

# the 11 contact masks with >= 2 stance feet (bit i = leg i in RL FL RR FR order)
MIXED_MASKS = np.array([m for m in range(16) if bin(m).count("1") >= 2], dtype=np.uint8)
assert len(MIXED_MASKS) == 11


def forward_kinematics(q, params: Params = None):
    """Body-frame foot positions, (..., 12) joint angles -> (..., 12); kinematics.cpp:96-100."""
    params = params or default_params()
    q = np.asarray(q, dtype=np.float64)
    hip = np.array(params.hip_offset[:]).reshape(4, 3)
    link = np.array(params.link[:]).reshape(4, 3)
    qq = q.reshape(q.shape[:-1] + (4, 3))
    t1, t2, t3 = qq[..., 0], qq[..., 1], qq[..., 2]
    l1, l2, l3 = link[:, 0], link[:, 1], link[:, 2]
    out = np.empty_like(qq)
    out[..., 0] = l2 * np.sin(t2) + l3 * np.sin(t2 + t3) + hip[:, 0]
    out[..., 1] = l1 * np.cos(t1) - l2 * np.sin(t1) * np.cos(t2) - l3 * np.sin(t1) * np.cos(t2 + t3) + hip[:, 1]
    out[..., 2] = l1 * np.sin(t1) + l2 * np.cos(t1) * np.cos(t2) + l3 * np.cos(t1) * np.cos(t2 + t3) + hip[:, 2]
    return out.reshape(q.shape)


def _rpy_matrix(roll, pitch, yaw):
    """R = Rz(yaw) Ry(pitch) Rx(roll), batched -> (n, 3, 3)."""
    cr, sr, cp, sp, cy, sy = np.cos(roll), np.sin(roll), np.cos(pitch), np.sin(pitch), np.cos(yaw), np.sin(yaw)
    R = np.empty(roll.shape + (3, 3))
    R[:, 0, 0] = cy * cp
    R[:, 0, 1] = cy * sp * sr - sy * cr
    R[:, 0, 2] = cy * sp * cr + sy * sr
    R[:, 1, 0] = sy * cp
    R[:, 1, 1] = sy * sp * sr + cy * cr
    R[:, 1, 2] = sy * sp * cr - cy * sr
    R[:, 2, 0] = -sp
    R[:, 2, 1] = cp * sr
    R[:, 2, 2] = cp * cr
    return R


def _exp_so3(v):
    """Rodrigues, (n,3) -> (n,3,3)."""
    th = np.linalg.norm(v, axis=1)
    small = th < 1e-12
    ths = np.where(small, 1.0, th)
    k = v / ths[:, None]
    K = np.zeros(v.shape[:1] + (3, 3))
    K[:, 0, 1], K[:, 0, 2] = -k[:, 2], k[:, 1]
    K[:, 1, 0], K[:, 1, 2] = k[:, 2], -k[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -k[:, 1], k[:, 0]
    s, c = np.sin(th)[:, None, None], np.cos(th)[:, None, None]
    E = np.eye(3)[None] + s * K + (1.0 - c) * (K @ K)
    E[small] = np.eye(3)
    return E


def stance_state(params: Params = None):
    """BASELINE config 1: one robot, 4-foot stance pose."""
    params = params or default_params()
    s = np.zeros(1, dtype=STATE_DTYPE)
    s["Rwb"][0] = np.eye(3).ravel()
    s["Rwb_d"][0] = np.eye(3).ravel()
    s["x"][0] = (0.0, 0.0, 0.2429)
    s["x_d"][0] = (0.0, 0.0, 0.26)  # commander_node.cpp:354
    s["q"][0] = STANCE_Q
    s["feet"][0] = forward_kinematics(STANCE_Q, params)
    s["contact"][0] = 1
    return s


def generate_states(n, seed, lo=0, profile="default", masks="all4", params: Params = None):
    """Records [lo, lo+n) of stream ``seed`` as a STATE_DTYPE array."""
    params = params or default_params()
    prof = PROFILES[profile] if isinstance(profile, str) else profile
    bg = np.random.Philox(key=int(seed))
    bg.advance(int(lo) * (DRAWS_PER_RECORD // 4))
    raw = bg.random_raw(int(n) * DRAWS_PER_RECORD).reshape(int(n), DRAWS_PER_RECORD)
    u = (raw >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)  # [0,1)
    del raw
    uni = u[:, :32]
    # Box-Muller on the upper half: 16 pairs -> 32 standard normals
    rad = np.sqrt(-2.0 * np.log(1.0 - u[:, 32:48]))
    ang = 2.0 * np.pi * u[:, 48:64]
    nrm = np.concatenate([rad * np.cos(ang), rad * np.sin(ang)], axis=1)

    def sym(col, half):  # U(-half, half)
        return (2.0 * uni[:, col] - 1.0) * half

    s = np.zeros(int(n), dtype=STATE_DTYPE)
    R = _rpy_matrix(sym(0, 0.3), sym(1, 0.3), sym(2, np.pi))
    direction = nrm[:, 0:3]
    dn = np.linalg.norm(direction, axis=1, keepdims=True)
    direction = direction / np.where(dn > 0, dn, 1.0)
    delta = direction * (uni[:, 3:4] * prof["att"])
    Rd = _exp_so3(delta) @ R
    s["Rwb"] = R.reshape(-1, 9)
    s["Rwb_d"] = Rd.reshape(-1, 9)
    x = np.array([0.0, 0.0, 0.26]) + (2.0 * uni[:, 4:7] - 1.0) * prof["pos"]
    s["x"] = x
    s["x_d"] = x + (2.0 * uni[:, 7:10] - 1.0) * prof["pos"]
    xdot = nrm[:, 3:6] * prof["vel"]
    s["xdot"] = xdot
    s["xdot_d"] = xdot + nrm[:, 6:9] * prof["vel_d"]
    w = nrm[:, 9:12] * prof["w"]
    s["w"] = w
    s["w_d"] = w + nrm[:, 12:15] * prof["w_d"]
    q = STANCE_Q + (2.0 * uni[:, 10:22] - 1.0) * 0.3
    s["q"] = q
    s["feet"] = forward_kinematics(q, params)
    if masks == "all4":
        s["contact"] = 1
    elif masks == "mixed":
        m = MIXED_MASKS[np.minimum((uni[:, 22] * 11).astype(np.int64), 10)]
        s["contact"] = (m[:, None] >> np.arange(4, dtype=np.uint8)) & 1
    else:
        raise ValueError(masks)
    return s


# BASELINE.json configs
CONFIGS = {
    "cfg1_stance": dict(n=1, seed=None, mu=0.8, masks="all4"),
    "cfg2_all4_65536": dict(n=65536, seed=20260102, mu=0.6, masks="all4"),
    "cfg3_mixed_1m": dict(n=1048576, seed=20260103, mu=0.6, masks="mixed"),
    "cfg5_mixed_8m": dict(n=8388608, seed=20260105, mu=0.6, masks="mixed"),
}


def generate_swing(S, seed, params: Params = None, reach=0.25):
    """Swing-leg references for the states ``S``: for every leg a joint-space target
    q + U(-reach, reach), mapped through FK to a body-frame foot position p_b and to the world frame as
    p_w = R (p_b + x) -- the inverse of the reference's own transform ``Rwb' * p_w - x`` (commander_node.cpp:491) --
    plus N(0, 0.5 m/s) reference velocities and N(0, 2 rad/s) measured joint velocities."""
    from .records import SWING_DTYPE

    params = params or default_params()
    rng = np.random.Generator(np.random.Philox(key=int(seed)))
    n = len(S)
    sw = np.zeros(n, dtype=SWING_DTYPE)
    q_ref = S["q"] + rng.uniform(-reach, reach, size=(n, 12))
    pb = forward_kinematics(q_ref, params).reshape(n, 4, 3)
    R = S["Rwb"].reshape(n, 3, 3)
    sw["foot_ref_pos"] = np.einsum("nij,nlj->nli", R, pb + S["x"][:, None, :]).reshape(n, 12)
    sw["foot_ref_vel"] = rng.normal(0.0, 0.5, size=(n, 12))
    sw["qdot"] = rng.normal(0.0, 2.0, size=(n, 12))
    return sw


# ---- BASELINE config 4: 10-step convex-MPC records ----------------------------
MPC_DRAWS_PER_RECORD = 32
CONFIGS["cfg4_mpc_65536"] = dict(n=65536, seed=20260104, mu=0.6)


def generate_mpc(n, seed, lo=0, params=None, gaits="mixed", scale=1.0):
    """Records [lo, lo+n) of MPC stream ``seed`` as an MPC_REC_DTYPE array (index-addressable like generate_states).

    Current state: roll, pitch ~ U(+-0.15), yaw ~ U(+-pi), p = (U(+-0.05), U(+-0.05), 0.26 + U(+-0.03)),
    omega ~ N(0, 0.5^2), v ~ N(0, 0.3^2).  Reference: level body at z = 0.26 moving with a commanded yaw-frame
    velocity (U(+-0.5), U(+-0.2), 0) and yaw rate U(+-0.5), integrated over the horizon.  Gait: stand (4 feet),
    trot (diagonal pairs alternating every 5 steps) or crawl (one leg swinging at a time), uniform, random phase.
    Feet: nominal stance rectangle (+-0.196, +-0.14) in the yaw frame + U(+-0.03), fixed in the world; lever arms
    are taken from the reference CoM of each step.  ``scale`` multiplies the attitude, twist and command magnitudes
    (a stress knob for the parity tests: more pyramid rows become active)."""
    from .records import MPC_REC_DTYPE, default_mpc_params

    params = params or default_mpc_params()
    n = int(n)
    bg = np.random.Philox(key=int(seed))
    bg.advance(int(lo) * (MPC_DRAWS_PER_RECORD // 4))
    raw = bg.random_raw(n * MPC_DRAWS_PER_RECORD).reshape(n, MPC_DRAWS_PER_RECORD)
    u = (raw >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    uni = u[:, :24]
    rad = np.sqrt(-2.0 * np.log(1.0 - u[:, 24:28]))
    ang = 2.0 * np.pi * u[:, 28:32]
    nrm = np.concatenate([rad * np.cos(ang), rad * np.sin(ang)], axis=1)  # 8 normals

    def sym(col, half):
        return (2.0 * uni[:, col] - 1.0) * half

    rec = np.zeros(n, dtype=MPC_REC_DTYPE)
    yaw0 = sym(2, np.pi)
    p0 = np.stack([sym(3, 0.05), sym(4, 0.05), 0.26 + sym(5, 0.03)], axis=1)
    rec["x0"][:, 0], rec["x0"][:, 1], rec["x0"][:, 2] = sym(0, 0.15 * scale), sym(1, 0.15 * scale), yaw0
    rec["x0"][:, 3:6] = p0
    rec["x0"][:, 6:9] = nrm[:, 0:3] * 0.5 * scale
    rec["x0"][:, 9:12] = nrm[:, 3:6] * 0.3 * scale
    rec["x0"][:, 12] = -9.81
    vcmd = np.stack([sym(6, 0.5 * scale), sym(7, 0.2 * scale)], axis=1)
    yawrate = sym(8, 0.5 * scale)
    c0, s0 = np.cos(yaw0), np.sin(yaw0)
    vw = np.stack([c0 * vcmd[:, 0] - s0 * vcmd[:, 1], s0 * vcmd[:, 0] + c0 * vcmd[:, 1]], axis=1)
    dt = params.dt
    steps = np.arange(1, 11, dtype=np.float64)
    xr = rec["xref"]
    xr[:, :, 2] = yaw0[:, None] + yawrate[:, None] * dt * steps
    xr[:, :, 3] = p0[:, 0:1] + vw[:, 0:1] * dt * steps
    xr[:, :, 4] = p0[:, 1:2] + vw[:, 1:2] * dt * steps
    xr[:, :, 5] = 0.26
    xr[:, :, 8] = yawrate[:, None]
    xr[:, :, 9] = vw[:, 0:1]
    xr[:, :, 10] = vw[:, 1:2]
    xr[:, :, 12] = -9.81
    # feet, fixed in the world
    sx = np.array([-1.0, 1.0, -1.0, 1.0]) * 0.196
    sy = np.array([1.0, 1.0, -1.0, -1.0]) * 0.14
    fx = sx[None, :] + (2.0 * uni[:, 9:13] - 1.0) * 0.03
    fy = sy[None, :] + (2.0 * uni[:, 13:17] - 1.0) * 0.03
    foot_w = np.empty((n, 4, 3))
    foot_w[:, :, 0] = p0[:, 0:1] + c0[:, None] * fx - s0[:, None] * fy
    foot_w[:, :, 1] = p0[:, 1:2] + s0[:, None] * fx + c0[:, None] * fy
    foot_w[:, :, 2] = 0.0
    rec["r"] = foot_w[:, None, :, :] - xr[:, :, None, 3:6]
    # gait
    k = np.arange(10)[None, :]
    phase = np.minimum((uni[:, 18] * 10).astype(np.int64), 9)[:, None]
    if gaits == "mixed":
        gait = np.minimum((uni[:, 17] * 3).astype(np.int64), 2)
    else:
        gait = np.full(n, {"stand": 0, "trot": 1, "crawl": 2}[gaits], dtype=np.int64)
    contact = np.ones((n, 10, 4), dtype=np.uint8)
    first = ((k + phase) % 10) < 5  # trot: RL+FR stance on the first half-period, FL+RR on the second
    trot = np.stack([first, ~first, ~first, first], axis=2).astype(np.uint8)
    swing_leg = ((k + phase) % 8) // 2  # crawl: leg swing_leg is in the air
    crawl = (np.arange(4)[None, None, :] != swing_leg[:, :, None]).astype(np.uint8)
    contact = np.where((gait == 1)[:, None, None], trot, contact)
    contact = np.where((gait == 2)[:, None, None], crawl, contact)
    rec["contact"] = contact
    return rec
