"""Flat record layouts of the batched balance-controller boundary.

These mirror ``include/qpb200.h`` byte for byte.  One ``STATE_DTYPE`` item is the argument
list of ``BalanceController::control`` (reference balance_controller.hpp:104-107) plus the joint
angles consumed by ``QuadrupedKinematics::jacobianTransposeControl`` (kinematics.hpp:106-107);
one ``OUT_DTYPE`` item is the ``ForceMap`` + ``TorqueMap`` those two calls return, flattened in
the reference's leg order RL, FL, RR, FR (commander_node.cpp:61).
"""
import ctypes

import numpy as np

LEG_NAMES = ("RL", "FL", "RR", "FR")

STATE_DTYPE = np.dtype(
    [
        ("Rwb", "<f8", (9,)),
        ("Rwb_d", "<f8", (9,)),
        ("x", "<f8", (3,)),
        ("xdot", "<f8", (3,)),
        ("w", "<f8", (3,)),
        ("x_d", "<f8", (3,)),
        ("xdot_d", "<f8", (3,)),
        ("w_d", "<f8", (3,)),
        ("feet", "<f8", (12,)),
        ("q", "<f8", (12,)),
        ("contact", "u1", (4,)),
        ("pad", "u1", (28,)),
    ]
)
assert STATE_DTYPE.itemsize == 512

OUT_DTYPE = np.dtype(
    [
        ("grf_body", "<f8", (12,)),
        ("tau", "<f8", (12,)),
        ("status", "<i4"),
        ("iters", "<i4"),
        ("pad", "u1", (56,)),
    ]
)
assert OUT_DTYPE.itemsize == 256

# wire records of the host-buffer calls (qpb_wire_state / qpb_wire_out): the same fields without the padding
WIRE_STATE_DTYPE = np.dtype(STATE_DTYPE.descr[:-1] + [("warm", "<u4")])
assert WIRE_STATE_DTYPE.itemsize == 488
WIRE_OUT_DTYPE = np.dtype([("grf_body", "<f8", (12,)), ("tau", "<f8", (12,)), ("status", "<i2"), ("iters", "<i2"), ("wset", "<u4")])
assert WIRE_OUT_DTYPE.itemsize == 200


def to_wire(states: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
    """qpb_state_rec array -> qpb_wire_state array (pad[0:4], the warm-start word, becomes ``warm``)."""
    w = np.empty(states.shape[0], dtype=WIRE_STATE_DTYPE) if out is None else out
    for name in STATE_DTYPE.names[:-1]:
        w[name] = states[name]
    w["warm"] = np.ascontiguousarray(states["pad"][:, :4]).view("<u4")[:, 0]
    return w


def from_wire(wout: np.ndarray) -> np.ndarray:
    """qpb_wire_out array -> qpb_out_rec array (for comparisons with the device-record entry points)."""
    o = np.zeros(wout.shape[0], dtype=OUT_DTYPE)
    o["grf_body"], o["tau"], o["status"], o["iters"] = wout["grf_body"], wout["tau"], wout["status"], wout["iters"]
    o["pad"][:, :4] = np.ascontiguousarray(wout["wset"]).view("u1").reshape(-1, 4)
    return o


# per-QP status (qpb200.h)
QPB_OK, QPB_MAX_ITER, QPB_BAD_INPUT = 0, 1, 2

# Algorithmic bytes per QP: 60 doubles + 4 contact bytes in, 24 doubles + status out.
ALGO_BYTES_PER_QP = 60 * 8 + 4 + 24 * 8 + 4


class Params(ctypes.Structure):
    """``qpb_params``: constructor arguments of BalanceController (balance_controller.hpp:85-88),
    the kinematic constants of QuadrupedKinematics (kinematics.cpp:23-47) and the caller's torque
    clamp (commander_node.cpp:324-325, 526).  Matrices are row-major."""

    _fields_ = [
        ("mu", ctypes.c_double),
        ("mass", ctypes.c_double),
        ("fzmin", ctypes.c_double),
        ("fzmax", ctypes.c_double),
        ("Ib", ctypes.c_double * 9),
        ("S", ctypes.c_double * 36),
        ("W", ctypes.c_double * 144),
        ("kff", ctypes.c_double * 6),
        ("kp_p", ctypes.c_double * 3),
        ("kd_p", ctypes.c_double * 3),
        ("kp_w", ctypes.c_double * 3),
        ("kd_w", ctypes.c_double * 3),
        ("hip_offset", ctypes.c_double * 12),
        ("link", ctypes.c_double * 12),
        ("tau_min", ctypes.c_double),
        ("tau_max", ctypes.c_double),
        ("clamp_tau", ctypes.c_int32),
        ("max_iter", ctypes.c_int32),
    ]

    def copy(self):
        other = Params()
        ctypes.memmove(ctypes.byref(other), ctypes.byref(self), ctypes.sizeof(self))
        return other


def default_params(mu=0.8):
    """Values of quadruped_simulation/config/mit_cheetah_config.yaml:66-99 with W = 1e-5*I
    (commander_node.cpp:289, 305) and the geometry of kinematics.cpp:23-47."""
    p = Params()
    p.mu, p.mass, p.fzmin, p.fzmax = mu, 11.0, 10.0, 120.0
    Ib = np.diag([0.011253, 0.036203, 0.042673])
    S = np.diag([1.0, 1.0, 1.0, 10.0, 10.0, 5.0])
    W = 1e-5 * np.eye(12)
    p.Ib[:] = Ib.ravel().tolist()
    p.S[:] = S.ravel().tolist()
    p.W[:] = W.ravel().tolist()
    p.kff[:] = [0.0, 0.0, 0.15, 0.0, 0.0, 0.0]
    p.kp_p[:] = [100.0] * 3
    p.kd_p[:] = [50.0] * 3
    p.kp_w[:] = [5000.0] * 3
    p.kd_w[:] = [500.0] * 3
    xbh, ybh, zbh = 0.196, 0.050, 0.0
    l1, l2, l3 = 0.077, 0.211, 0.230
    sx = (-1.0, 1.0, -1.0, 1.0)  # RL FL RR FR
    sy = (1.0, 1.0, -1.0, -1.0)
    hip, link = [], []
    for leg in range(4):
        hip += [sx[leg] * xbh, sy[leg] * ybh, zbh]
        link += [sy[leg] * l1, -l2, -l3]
    p.hip_offset[:] = hip
    p.link[:] = link
    p.tau_min, p.tau_max, p.clamp_tau, p.max_iter = -20.0, 20.0, 0, 200
    return p


# ---- swing-leg half of the control tick ----------------------------------------
# One item per robot: the reference foot states FootTrajectoryManager::referenceState returns for swing legs
# (world frame, commander_node.cpp:487-488) and the measured joint velocities (JointStatesMap.qdot).
SWING_DTYPE = np.dtype([("foot_ref_pos", "<f8", (12,)), ("foot_ref_vel", "<f8", (12,)), ("qdot", "<f8", (12,))])
assert SWING_DTYPE.itemsize == 288


class JointGains(ctypes.Structure):
    """``qpb_joint_gains``: JointController(kff, kp, kd) (joint_controller.hpp; commander_node.cpp:341)."""

    _fields_ = [("kff", ctypes.c_double * 3), ("kp", ctypes.c_double * 3), ("kd", ctypes.c_double * 3)]


def default_joint_gains():
    """joint_control/{kff,kp,kd} of mit_cheetah_config.yaml:50-53."""
    g = JointGains()
    g.kff[:] = [0.0, 0.0, 0.0]
    g.kp[:] = [40.0, 40.0, 50.0]
    g.kd[:] = [1.0, 1.0, 1.0]
    return g


# ---- 10-step convex-MPC QP (BASELINE config 4) -----------------------------------
# The reference has no code for this path; the formulation is stated in include/qpb200.h (qpb_mpc_*).
MPC_HORIZON, MPC_NV, MPC_NC = 10, 120, 200

MPC_REC_DTYPE = np.dtype(
    [
        ("x0", "<f8", (13,)),
        ("xref", "<f8", (10, 13)),
        ("r", "<f8", (10, 4, 3)),
        ("contact", "u1", (10, 4)),
        ("pad", "u1", (32,)),
    ]
)
assert MPC_REC_DTYPE.itemsize == 2176

MPC_OUT_DTYPE = np.dtype([("U", "<f8", (120,)), ("status", "<i4"), ("iters", "<i4"), ("pad", "u1", (56,))])
assert MPC_OUT_DTYPE.itemsize == 1024

# Algorithmic bytes per MPC QP: 263 doubles + 40 contact bytes in, 120 doubles + status out.
MPC_ALGO_BYTES_PER_QP = 263 * 8 + 40 + 120 * 8 + 4


class MpcParams(ctypes.Structure):
    """``qpb_mpc_params``."""

    _fields_ = [
        ("mu", ctypes.c_double),
        ("mass", ctypes.c_double),
        ("fzmin", ctypes.c_double),
        ("fzmax", ctypes.c_double),
        ("Ib", ctypes.c_double * 9),
        ("dt", ctypes.c_double),
        ("Lw", ctypes.c_double * 13),
        ("alpha", ctypes.c_double),
        ("max_iter", ctypes.c_int32),
        ("pad", ctypes.c_int32),
    ]

    def copy(self):
        other = MpcParams()
        ctypes.memmove(ctypes.byref(other), ctypes.byref(self), ctypes.sizeof(self))
        return other


def default_mpc_params(mu=0.6):
    """Robot constants of mit_cheetah_config.yaml:95-99; horizon step and weights of Di Carlo et al. (IROS 2018)."""
    p = MpcParams()
    p.mu, p.mass, p.fzmin, p.fzmax = mu, 11.0, 10.0, 120.0
    p.Ib[:] = np.diag([0.011253, 0.036203, 0.042673]).ravel().tolist()
    p.dt = 0.03
    p.Lw[:] = [0.25, 0.25, 10.0, 2.0, 2.0, 50.0, 0.0, 0.0, 0.3, 0.2, 0.2, 0.1, 0.0]
    p.alpha = 4e-5
    p.max_iter = 1000
    return p


# ---- foothold planner + swing-foot trajectory and message adapters (rank 3) --------------
# One PLAN item per robot: the reference's FootTrajBounds per leg (types.hpp:52-65; the state FootTrajectoryManager
# keeps in traj_map_), the gait phase of every leg (GaitMap[leg].second) and the "switched stance -> swing this tick"
# flags FootPlanner::updateStates derives (foot_planner.cpp:106-157).
PLAN_DTYPE = np.dtype([("p_start", "<f8", (12,)), ("p_final", "<f8", (12,)), ("phase", "<f8", (4,)), ("replan", "u1", (4,)),
                       ("pad", "u1", (12,))])
assert PLAN_DTYPE.itemsize == 240

# quadruped_msgs/CoMState in message field order (pose.position, pose.orientation x y z w, twist.linear, twist.angular)
COM_MSG_DTYPE = np.dtype([("position", "<f8", (3,)), ("orientation", "<f8", (4,)), ("linear", "<f8", (3,)), ("angular", "<f8", (3,))])
assert COM_MSG_DTYPE.itemsize == 104
# sensor_msgs/JointState position / velocity in joint_names order (mit_cheetah_config.yaml:35-37)
JOINT_MSG_DTYPE = np.dtype([("position", "<f8", (12,)), ("velocity", "<f8", (12,))])
assert JOINT_MSG_DTYPE.itemsize == 192
# quadruped_msgs/JointTorqueCmd.torque as the reference fills it (legs in std::map order FL FR RL RR), entry count, and
# the leg each entry belongs to (stands in for actuator_name, commander_node.cpp:517-533)
TORQUE_CMD_DTYPE = np.dtype([("torque", "<f8", (12,)), ("leg", "u1", (12,)), ("count", "<i4")])
assert TORQUE_CMD_DTYPE.itemsize == 112


class PlanParams(ctypes.Structure):
    """``qpb_plan_params``: FootPlanner constants (foot_planner.cpp:22-42) and the gait/ parameters
    FootTrajectoryManager is constructed with (commander_node.cpp:245-247, 360-361)."""

    _fields_ = [
        ("k_raibert", ctypes.c_double),
        ("g", ctypes.c_double),
        ("thigh_offset", ctypes.c_double * 12),
        ("height", ctypes.c_double),
        ("t_swing", ctypes.c_double),
        ("t_stance", ctypes.c_double),
    ]


def default_plan_params():
    p = PlanParams()
    p.k_raibert, p.g = 0.01, 9.81
    sx, sy = (-1.0, 1.0, -1.0, 1.0), (1.0, 1.0, -1.0, -1.0)
    off = []
    for leg in range(4):
        off += [sx[leg] * 0.196, sy[leg] * 0.127, 0.0]
    p.thigh_offset[:] = off
    p.height, p.t_swing, p.t_stance = 0.08, 0.18, 0.8  # mit_cheetah_config.yaml:17-19
    return p
