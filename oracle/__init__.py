"""ctypes binding of the CPU oracle.  TEST INFRASTRUCTURE ONLY (see qpb_oracle.h):
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
import ctypes
import os
import subprocess

import numpy as np

from quadruped_control_b200.records import OUT_DTYPE, STATE_DTYPE, SWING_DTYPE, JointGains, Params

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libqpb_oracle.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < max(
        os.path.getmtime(os.path.join(_HERE, f)) for f in ("qpb_oracle.c", "qpb_oracle.h", "mpc_oracle.c", "mpc_oracle.h", "plan_oracle.c", "plan_oracle.h")
    ):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        vp = ctypes.c_void_p
        L.orc_default_params.argtypes = [vp]
        L.orc_forward_kinematics.argtypes = [vp, ctypes.c_int, dp, dp]
        L.orc_leg_jacobian.argtypes = [vp, ctypes.c_int, dp, dp]
        L.orc_angle_axis_total.argtypes = [dp, dp]
        L.orc_assemble.argtypes = [vp, vp, dp, dp, dp, dp, dp, dp, dp]
        L.orc_qp_solve.argtypes = [ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, dp, ctypes.c_int, dp, dp,
                                   ctypes.POINTER(ctypes.c_int)]
        L.orc_qp_solve.restype = ctypes.c_int
        L.orc_control.argtypes = [vp, vp, vp, dp]
        L.orc_control.restype = ctypes.c_int
        L.orc_control_batch.argtypes = [vp, vp, ctypes.c_int64, vp, ctypes.c_int]
        L.orc_default_joint_gains.argtypes = [vp]
        L.orc_leg_inverse_kinematics.argtypes = [vp, ctypes.c_int, dp, dp]
        L.orc_leg_jacobian_inverse.argtypes = [vp, ctypes.c_int, dp, dp]
        L.orc_leg_jacobian_inverse.restype = ctypes.c_int
        L.orc_tick_batch.argtypes = [vp, vp, vp, vp, ctypes.c_int64, vp, ctypes.c_int]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def default_params():
    p = Params()  # same byte layout as orc_params
    lib().orc_default_params(ctypes.byref(p))
    return p


def forward_kinematics(params, leg, q):
    q = np.ascontiguousarray(q, dtype=np.float64)
    out = np.empty(3)
    lib().orc_forward_kinematics(ctypes.byref(params), leg, _dp(q), _dp(out))
    return out


def leg_jacobian(params, leg, q):
    q = np.ascontiguousarray(q, dtype=np.float64)
    out = np.empty(9)
    lib().orc_leg_jacobian(ctypes.byref(params), leg, _dp(q), _dp(out))
    return out.reshape(3, 3)


def angle_axis_total(R):
    R = np.ascontiguousarray(R, dtype=np.float64).reshape(9)
    out = np.empty(3)
    lib().orc_angle_axis_total(_dp(R), _dp(out))
    return out


def assemble(params, state):
    """QP of one STATE_DTYPE record in the reference's qpOASES form -> dict(Q,c,C,lb,ub,A,b)."""
    state = np.ascontiguousarray(state).reshape(1)
    assert state.dtype == STATE_DTYPE
    Q, c, C = np.empty(144), np.empty(12), np.empty(240)
    lb, ub, A, b = np.empty(20), np.empty(20), np.empty(72), np.empty(6)
    lib().orc_assemble(ctypes.byref(params), state.ctypes.data, _dp(Q), _dp(c), _dp(C), _dp(lb), _dp(ub), _dp(A), _dp(b))
    return dict(Q=Q.reshape(12, 12), c=c, C=C.reshape(20, 12), lb=lb, ub=ub, A=A.reshape(6, 12), b=b)


def qp_solve(Q, c, C, lb, ub, max_iter=200):
    Q = np.ascontiguousarray(Q, dtype=np.float64)
    C = np.ascontiguousarray(C, dtype=np.float64)
    c, lb, ub = (np.ascontiguousarray(v, dtype=np.float64) for v in (c, lb, ub))
    n, m = Q.shape[0], C.shape[0]
    x, lam = np.empty(n), np.empty(m)
    it = ctypes.c_int(0)
    st = lib().orc_qp_solve(n, m, _dp(Q), _dp(c), _dp(C), _dp(lb), _dp(ub), max_iter, _dp(x), _dp(lam), ctypes.byref(it))
    return st, x, lam, it.value


def control(params, state):
    """One record -> (OUT_DTYPE scalar array of 1, world-frame QP solution fw)."""
    state = np.ascontiguousarray(state).reshape(1)
    out = np.zeros(1, dtype=OUT_DTYPE)
    fw = np.zeros(12)
    lib().orc_control(ctypes.byref(params), state.ctypes.data, out.ctypes.data, _dp(fw))
    return out, fw


def control_batch(params, states, nthreads=1):
    states = np.ascontiguousarray(states)
    assert states.dtype == STATE_DTYPE
    out = np.zeros(states.shape[0], dtype=OUT_DTYPE)
    lib().orc_control_batch(ctypes.byref(params), states.ctypes.data, states.shape[0], out.ctypes.data, int(nthreads))
    return out


# ---- oracle/_ref: the reference's own sources compiled against stand-in third-party headers ------
_REF_PATH = os.path.join(_HERE, "_ref", "libqpb_ref.so")
_ref = None


def ref_build():
    """(Re)build oracle/_ref when /root/reference is present; elsewhere the prebuilt library is used."""
    subprocess.run(["bash", os.path.join(_HERE, "ref_build.sh")], check=True, capture_output=True)
    return os.path.exists(_REF_PATH)


def ref_available():
    return os.path.exists(_REF_PATH) or (os.path.isdir("/root/reference") and ref_build())


def ref_lib():
    global _ref
    if _ref is None:
        if os.path.isdir("/root/reference"):
            ref_build()
        R = ctypes.CDLL(_REF_PATH)
        dp = ctypes.POINTER(ctypes.c_double)
        R.ref_control.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        R.ref_control.restype = ctypes.c_int
        R.ref_control_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_int]
        R.ref_forward_kinematics.argtypes = [dp, dp]
        R.ref_leg_jacobian.argtypes = [ctypes.c_int, dp, dp]
        R.ref_error_count.restype = ctypes.c_int
        R.ref_leg_inverse_kinematics.argtypes = [ctypes.c_int, dp, dp]
        R.ref_leg_jacobian_inverse.argtypes = [ctypes.c_int, dp, dp]
        R.ref_tick_batch.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_longlong, ctypes.c_void_p, ctypes.c_int]
        _ref = R
    return _ref


def ref_control_batch(params, states, nthreads=1):
    """States through the reference's BalanceController::control + jacobianTransposeControl
    (its own source, stand-in Armadillo/qpOASES/ROS/rigid3d; one controller instance per thread)."""
    states = np.ascontiguousarray(states)
    assert states.dtype == STATE_DTYPE
    out = np.zeros(states.shape[0], dtype=OUT_DTYPE)
    ref_lib().ref_control_batch(ctypes.byref(params), states.ctypes.data, states.shape[0], out.ctypes.data, int(nthreads))
    return out


def ref_forward_kinematics(q12):
    q = np.ascontiguousarray(q12, dtype=np.float64).reshape(12)
    out = np.empty(12)
    ref_lib().ref_forward_kinematics(_dp(q), _dp(out))
    return out


def ref_leg_jacobian(leg, q):
    q = np.ascontiguousarray(q, dtype=np.float64).reshape(3)
    out = np.empty(9)
    ref_lib().ref_leg_jacobian(int(leg), _dp(q), _dp(out))
    return out.reshape(3, 3)


# ---- swing-leg half of the control tick -------------------------------------------
def default_joint_gains():
    g = JointGains()
    lib().orc_default_joint_gains(ctypes.byref(g))
    return g


def leg_inverse_kinematics(params, leg, foothold):
    f = np.ascontiguousarray(foothold, dtype=np.float64).reshape(3)
    out = np.empty(3)
    lib().orc_leg_inverse_kinematics(ctypes.byref(params), int(leg), _dp(f), _dp(out))
    return out


def leg_jacobian_inverse(params, leg, q):
    q = np.ascontiguousarray(q, dtype=np.float64).reshape(3)
    out = np.empty(9)
    kind = lib().orc_leg_jacobian_inverse(ctypes.byref(params), int(leg), _dp(q), _dp(out))
    return out.reshape(3, 3), kind


def tick_batch(params, gains, states, swing, nthreads=1):
    """control() + jacobianTransposeControl() + swing-leg joint PD, merged and clamped (commander_node.cpp:482-533)."""
    states, swing = np.ascontiguousarray(states), np.ascontiguousarray(swing)
    assert states.dtype == STATE_DTYPE and swing.dtype == SWING_DTYPE and len(states) == len(swing)
    out = np.zeros(len(states), dtype=OUT_DTYPE)
    lib().orc_tick_batch(ctypes.byref(params), ctypes.byref(gains), states.ctypes.data, swing.ctypes.data, len(states),
                         out.ctypes.data, int(nthreads))
    return out


def ref_leg_inverse_kinematics(leg, foothold):
    f = np.ascontiguousarray(foothold, dtype=np.float64).reshape(3)
    out = np.empty(3)
    ref_lib().ref_leg_inverse_kinematics(int(leg), _dp(f), _dp(out))
    return out


def ref_leg_jacobian_inverse(leg, q):
    q = np.ascontiguousarray(q, dtype=np.float64).reshape(3)
    out = np.empty(9)
    ref_lib().ref_leg_jacobian_inverse(int(leg), _dp(q), _dp(out))
    return out.reshape(3, 3)


def ref_tick_batch(params, gains, states, swing, nthreads=1):
    states, swing = np.ascontiguousarray(states), np.ascontiguousarray(swing)
    assert states.dtype == STATE_DTYPE and swing.dtype == SWING_DTYPE and len(states) == len(swing)
    out = np.zeros(len(states), dtype=OUT_DTYPE)
    ref_lib().ref_tick_batch(ctypes.byref(params), ctypes.byref(gains), states.ctypes.data, swing.ctypes.data, len(states),
                             out.ctypes.data, int(nthreads))
    return out


# ---- 10-step convex-MPC QP (BASELINE config 4; parity unpinned: no reference code exists) ---------------------
def _mpc_protos():
    L = lib()
    if not getattr(L, "_mpc_ready", False):
        vp = ctypes.c_void_p
        dp = ctypes.POINTER(ctypes.c_double)
        L.orc_mpc_default_params.argtypes = [vp]
        L.orc_mpc_assemble.argtypes = [vp, vp, dp, dp, dp, dp, dp]
        L.orc_mpc_solve.argtypes = [vp, vp, vp]
        L.orc_mpc_solve.restype = ctypes.c_int
        L.orc_mpc_batch.argtypes = [vp, vp, ctypes.c_int64, vp, ctypes.c_int]
        L._mpc_ready = True
    return L


def mpc_default_params():
    from quadruped_control_b200.records import MpcParams

    p = MpcParams()
    _mpc_protos().orc_mpc_default_params(ctypes.byref(p))
    return p


def mpc_assemble(params, rec):
    from quadruped_control_b200.records import MPC_REC_DTYPE

    rec = np.ascontiguousarray(rec).reshape(1)
    assert rec.dtype == MPC_REC_DTYPE
    H, g, C = np.empty(120 * 120), np.empty(120), np.empty(200 * 120)
    lb, ub = np.empty(200), np.empty(200)
    _mpc_protos().orc_mpc_assemble(ctypes.byref(params), rec.ctypes.data, _dp(H), _dp(g), _dp(C), _dp(lb), _dp(ub))
    return dict(Q=H.reshape(120, 120), c=g, C=C.reshape(200, 120), lb=lb, ub=ub)


def mpc_batch(params, recs, nthreads=1):
    from quadruped_control_b200.records import MPC_OUT_DTYPE, MPC_REC_DTYPE

    recs = np.ascontiguousarray(recs)
    assert recs.dtype == MPC_REC_DTYPE
    out = np.zeros(len(recs), dtype=MPC_OUT_DTYPE)
    _mpc_protos().orc_mpc_batch(ctypes.byref(params), recs.ctypes.data, len(recs), out.ctypes.data, int(nthreads))
    return out


# ---- foothold planner + swing-foot trajectory, message adapters (rank 3) ---------------
def _plan_protos():
    L = lib()
    if not getattr(L, "_plan_ready", False):
        vp, dp = ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)
        L.orc_default_plan_params.argtypes = [vp]
        L.orc_single_foot.argtypes = [vp, ctypes.c_int, vp, dp]
        L.orc_foot_trajectory.argtypes = [dp, dp, dp, dp]
        L.orc_foot_trajectory.restype = ctypes.c_int
        L.orc_track_trajectory.argtypes = [dp, ctypes.c_double, dp, dp]
        L.orc_plan_batch.argtypes = [vp, vp, vp, vp, ctypes.c_int64]
        L.orc_adapt_inputs.argtypes = [vp, vp, vp, vp, vp]
        L.orc_torque_cmd.argtypes = [vp, dp, ctypes.POINTER(ctypes.c_int), dp, ctypes.POINTER(ctypes.c_int)]
        L.orc_torque_cmd.restype = ctypes.c_int
        L._plan_ready = True
    return L


def plan_default_params():
    from quadruped_control_b200.records import PlanParams

    p = PlanParams()
    _plan_protos().orc_default_plan_params(ctypes.byref(p))
    return p


def single_foot(pp, leg, state):
    state = np.ascontiguousarray(state).reshape(1)
    out = np.empty(3)
    _plan_protos().orc_single_foot(ctypes.byref(pp), int(leg), state.ctypes.data, _dp(out))
    return out


def foot_trajectory(p_start, p_center, p_final):
    a, b, c = (np.ascontiguousarray(v, dtype=np.float64) for v in (p_start, p_center, p_final))
    coef = np.empty(21)
    rc = _plan_protos().orc_foot_trajectory(_dp(a), _dp(b), _dp(c), _dp(coef))
    return rc, coef.reshape(7, 3)


def track_trajectory(coef, t):
    coef = np.ascontiguousarray(coef, dtype=np.float64).reshape(21)
    pos, vel = np.empty(3), np.empty(3)
    _plan_protos().orc_track_trajectory(_dp(coef), float(t), _dp(pos), _dp(vel))
    return pos, vel


def plan_batch(pp, states, plan, swing):
    """In place: re-plans the flagged swing legs in ``plan`` and writes the reference foot states into ``swing``."""
    from quadruped_control_b200.records import PLAN_DTYPE

    assert states.dtype == STATE_DTYPE and plan.dtype == PLAN_DTYPE and swing.dtype == SWING_DTYPE
    assert states.flags.c_contiguous and plan.flags.c_contiguous and swing.flags.c_contiguous
    _plan_protos().orc_plan_batch(ctypes.byref(pp), states.ctypes.data, plan.ctypes.data, swing.ctypes.data, len(states))


def ref_plan_batch(pp, states, plan, swing):
    """Same through the reference's own FootPlanner / FootTrajectoryManager (oracle/_ref)."""
    R = ref_lib()
    R.ref_plan_batch.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_longlong]
    R.ref_plan_batch(ctypes.byref(pp), states.ctypes.data, plan.ctypes.data, swing.ctypes.data, len(states))


def ref_two_tick_replan(pp, leg_a, a_start, a_final, phase_a, leg_b, b_start, b_final, phase_b):
    """Foot reference states of legs a and b after leg b re-planned while leg a was mid-swing, through ONE
    FootTrajectoryManager of the reference (oracle/_ref) -> (pos+vel of a, pos+vel of b)."""
    R = ref_lib()
    dp = ctypes.POINTER(ctypes.c_double)
    R.ref_two_tick_replan.argtypes = [ctypes.c_void_p, ctypes.c_int, dp, dp, ctypes.c_double, ctypes.c_int, dp, dp, ctypes.c_double, dp, dp]
    arrs = [np.ascontiguousarray(v, dtype=np.float64) for v in (a_start, a_final, b_start, b_final)]
    oa, ob = np.zeros(6), np.zeros(6)
    R.ref_two_tick_replan(ctypes.byref(pp), int(leg_a), _dp(arrs[0]), _dp(arrs[1]), float(phase_a), int(leg_b), _dp(arrs[2]),
                          _dp(arrs[3]), float(phase_b), _dp(oa), _dp(ob))
    return oa, ob


def ref_single_foot(t_stance, leg, state):
    state = np.ascontiguousarray(state).reshape(1)
    out = np.empty(3)
    R = ref_lib()
    R.ref_single_foot.argtypes = [ctypes.c_double, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)]
    R.ref_single_foot(float(t_stance), int(leg), state.ctypes.data, _dp(out))
    return out


def adapt_inputs(params, com, joints, states, swing):
    """In place: CoMState + JointState messages -> Rwb, x, xdot, w, q, feet of ``states`` and qdot of ``swing``."""
    L = _plan_protos()
    for i in range(len(com)):
        L.orc_adapt_inputs(ctypes.byref(params), com[i:i + 1].ctypes.data, joints[i:i + 1].ctypes.data,
                           states[i:i + 1].ctypes.data, swing[i:i + 1].ctypes.data)


def torque_cmd(params, tau, present):
    tau = np.ascontiguousarray(tau, dtype=np.float64).reshape(12)
    pres = (ctypes.c_int * 4)(*[int(v) for v in present])
    torque = np.zeros(12)
    legs = (ctypes.c_int * 12)()
    n = _plan_protos().orc_torque_cmd(ctypes.byref(params), _dp(tau), pres, _dp(torque), legs)
    return n, torque, np.array(list(legs), dtype=np.int64)
