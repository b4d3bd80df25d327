"""ctypes binding of libqpb200.so (include/qpb200.h).

This is plumbing for the tests and the benchmark; the product boundary is the C ABI itself and the
C++ shim in ``cpp/balance_controller.hpp``.  There is no CPU fallback: if the library is missing,
or no CUDA device is present, every compute call raises.
"""
import ctypes
import os

import numpy as np

from .records import (MPC_OUT_DTYPE, MPC_REC_DTYPE, OUT_DTYPE, STATE_DTYPE, SWING_DTYPE, WIRE_OUT_DTYPE, WIRE_STATE_DTYPE, JointGains, MpcParams, Params,
                      PlanParams)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QPB_LIB") or os.path.join(_HERE, "libqpb200.so")  # QPB_LIB: experiment builds

EXPORTS = (
    "qpb_version",
    "qpb_last_error",
    "qpb_default_params",
    "qpb_create",
    "qpb_destroy",
    "qpb_control_batch_packed",
    "qpb_control_batch",
    "qpb_control_batch_host",
    "qpb_control_batch_host_async",
    "qpb_host_sync",
    "qpb_control_batch_wire_host",
    "qpb_control_batch_wire_host_async",
    "qpb_jt_batch",
    "qpb_fk_batch",
    "qpb_jt_batch_host",
    "qpb_fk_batch_host",
    "qpb_set_warm_batches",
    "qpb_set_joint_gains",
    "qpb_tick_batch_packed",
    "qpb_tick_batch_host",
    "qpb_host_alloc",
    "qpb_host_free",
    "qpb_launch_count",
    "qpb_default_plan_params",
    "qpb_set_plan_params",
    "qpb_plan_batch",
    "qpb_adapt_inputs_batch",
    "qpb_torque_cmd_batch",
    "qpb_mpc_default_params",
    "qpb_mpc_create",
    "qpb_mpc_destroy",
    "qpb_mpc_batch_packed",
    "qpb_mpc_batch_host",
    "qpb_mpc_launch_count",
    "qpb_device_count",
    "qpb_multi_create",
    "qpb_multi_destroy",
    "qpb_multi_num_shards",
    "qpb_multi_shard_range",
    "qpb_multi_set_joint_gains",
    "qpb_multi_control_batch_host",
    "qpb_multi_tick_batch_host",
    "qpb_multi_launch_count",
)

_lib = None


class QpbError(RuntimeError):
    pass


def load():
    """dlopen libqpb200.so and declare its prototypes.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise QpbError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback)")
    L = ctypes.CDLL(LIB_PATH)
    vp, i64, dp = ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p
    L.qpb_version.restype = ctypes.c_int
    L.qpb_last_error.restype = ctypes.c_char_p
    L.qpb_default_params.argtypes = [ctypes.POINTER(Params)]
    L.qpb_create.argtypes = [ctypes.POINTER(Params), ctypes.c_int, ctypes.POINTER(vp)]
    L.qpb_destroy.argtypes = [vp]
    L.qpb_control_batch_packed.argtypes = [vp, i64, vp, vp, vp]
    L.qpb_control_batch.argtypes = [vp, i64] + [dp] * 14 + [vp]
    L.qpb_control_batch_host.argtypes = [vp, i64, vp, vp]
    L.qpb_control_batch_host_async.argtypes = [vp, i64, vp, vp]
    L.qpb_host_sync.argtypes = [vp]
    L.qpb_control_batch_wire_host.argtypes = [vp, i64, vp, vp]
    L.qpb_control_batch_wire_host_async.argtypes = [vp, i64, vp, vp]
    L.qpb_jt_batch.argtypes = [vp, i64, dp, dp, dp, dp, vp]
    L.qpb_fk_batch.argtypes = [vp, i64, dp, dp, vp]
    L.qpb_jt_batch_host.argtypes = [vp, i64, dp, dp, dp, dp]
    L.qpb_fk_batch_host.argtypes = [vp, i64, dp, dp]
    L.qpb_set_warm_batches.argtypes = [vp, ctypes.c_int]
    L.qpb_set_joint_gains.argtypes = [vp, ctypes.POINTER(JointGains)]
    L.qpb_tick_batch_packed.argtypes = [vp, i64, vp, vp, vp, vp]
    L.qpb_tick_batch_host.argtypes = [vp, i64, vp, vp, vp]
    L.qpb_host_alloc.argtypes = [ctypes.POINTER(vp), ctypes.c_size_t]
    L.qpb_host_free.argtypes = [vp]
    L.qpb_launch_count.argtypes = [vp]
    L.qpb_launch_count.restype = i64
    L.qpb_default_plan_params.argtypes = [ctypes.POINTER(PlanParams)]
    L.qpb_set_plan_params.argtypes = [vp, ctypes.POINTER(PlanParams)]
    L.qpb_plan_batch.argtypes = [vp, i64, vp, vp, vp, vp]
    L.qpb_adapt_inputs_batch.argtypes = [vp, i64, vp, vp, vp, vp, vp]
    L.qpb_torque_cmd_batch.argtypes = [vp, i64, vp, vp, vp, vp]
    L.qpb_mpc_default_params.argtypes = [ctypes.POINTER(MpcParams)]
    L.qpb_mpc_create.argtypes = [ctypes.POINTER(MpcParams), ctypes.c_int, ctypes.POINTER(vp)]
    L.qpb_mpc_destroy.argtypes = [vp]
    L.qpb_mpc_batch_packed.argtypes = [vp, i64, vp, vp, vp]
    L.qpb_mpc_batch_host.argtypes = [vp, i64, vp, vp]
    L.qpb_mpc_launch_count.argtypes = [vp]
    L.qpb_mpc_launch_count.restype = i64
    L.qpb_device_count.argtypes = [ctypes.POINTER(ctypes.c_int)]
    L.qpb_multi_create.argtypes = [ctypes.POINTER(Params), ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.POINTER(vp)]
    L.qpb_multi_destroy.argtypes = [vp]
    L.qpb_multi_num_shards.argtypes = [vp]
    L.qpb_multi_shard_range.argtypes = [i64, ctypes.c_int, ctypes.c_int, ctypes.POINTER(i64), ctypes.POINTER(i64)]
    L.qpb_multi_set_joint_gains.argtypes = [vp, ctypes.POINTER(JointGains)]
    L.qpb_multi_control_batch_host.argtypes = [vp, i64, vp, vp]
    L.qpb_multi_tick_batch_host.argtypes = [vp, i64, vp, vp, vp]
    L.qpb_multi_launch_count.argtypes = [vp]
    L.qpb_multi_launch_count.restype = i64
    for name in EXPORTS:
        getattr(L, name)
    _lib = L
    return L


def _check(rc, what):
    if rc != 0:
        raise QpbError(f"{what} failed ({rc}): {load().qpb_last_error().decode()}")


def default_params():
    p = Params()
    _check(load().qpb_default_params(ctypes.byref(p)), "qpb_default_params")
    return p


def _ptr(t):
    """Device/host address of a torch tensor, numpy array, or raw int."""
    if t is None:
        return None
    if isinstance(t, int):
        return t
    if isinstance(t, np.ndarray):
        return t.ctypes.data
    return t.data_ptr()


class PinnedBuffer:
    """Pinned host memory from qpb_host_alloc, viewed as a numpy array of ``dtype``."""

    def __init__(self, n, dtype):
        self.dtype = np.dtype(dtype)
        self.nbytes = int(n) * self.dtype.itemsize
        p = ctypes.c_void_p()
        _check(load().qpb_host_alloc(ctypes.byref(p), max(self.nbytes, 1)), "qpb_host_alloc")
        self.ptr = p.value
        buf = (ctypes.c_char * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(n))

    def free(self):
        if self.ptr:
            self.array = None
            load().qpb_host_free(self.ptr)
            self.ptr = None


def multi_shard_range(n, shard, num_shards):
    """[lo, hi) of ``shard`` as libqpb200's single-process multi-GPU calls cut the batch (no device needed)."""
    lo, hi = ctypes.c_int64(), ctypes.c_int64()
    _check(load().qpb_multi_shard_range(int(n), int(shard), int(num_shards), ctypes.byref(lo), ctypes.byref(hi)),
           "qpb_multi_shard_range")
    return lo.value, hi.value


class MultiBalanceSolver:
    """Owns one ``qpb_multi_handle``: the host-buffer calls sharded over several devices inside one process
    (one persistent host thread and one qpb_handle per device, no data-path collective)."""

    def __init__(self, params: Params = None, devices=None):
        L = load()
        self.params = params.copy() if params is not None else default_params()
        h = ctypes.c_void_p()
        if devices is None:
            arr, nd = None, 0
        else:
            nd = len(devices)
            arr = (ctypes.c_int * nd)(*[int(d) for d in devices])
        _check(L.qpb_multi_create(ctypes.byref(self.params), arr, nd, ctypes.byref(h)), "qpb_multi_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            load().qpb_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def num_shards(self):
        return int(load().qpb_multi_num_shards(self._h))

    @property
    def launches(self):
        return int(load().qpb_multi_launch_count(self._h))

    def set_joint_gains(self, gains: JointGains):
        _check(load().qpb_multi_set_joint_gains(self._h, ctypes.byref(gains)), "qpb_multi_set_joint_gains")

    def control_host(self, states: np.ndarray, out: np.ndarray = None):
        states = np.ascontiguousarray(states)
        assert states.dtype == STATE_DTYPE
        if out is None:
            out = np.empty(states.shape[0], dtype=OUT_DTYPE)
        assert out.dtype == OUT_DTYPE and out.shape[0] == states.shape[0] and out.flags.c_contiguous
        _check(load().qpb_multi_control_batch_host(self._h, states.shape[0], states.ctypes.data, out.ctypes.data),
               "qpb_multi_control_batch_host")
        return out

    def tick_host(self, states: np.ndarray, swing: np.ndarray, out: np.ndarray = None):
        states, swing = np.ascontiguousarray(states), np.ascontiguousarray(swing)
        assert states.dtype == STATE_DTYPE and swing.dtype == SWING_DTYPE and len(states) == len(swing)
        if out is None:
            out = np.empty(states.shape[0], dtype=OUT_DTYPE)
        assert out.dtype == OUT_DTYPE and out.shape[0] == states.shape[0] and out.flags.c_contiguous
        _check(load().qpb_multi_tick_batch_host(self._h, states.shape[0], states.ctypes.data, swing.ctypes.data,
                                                out.ctypes.data), "qpb_multi_tick_batch_host")
        return out


class BalanceSolver:
    """Owns one ``qpb_handle`` (replaces the BalanceController + QuadrupedKinematics pair the
    reference constructs at commander_node.cpp:337-338, 358) on one CUDA device."""

    def __init__(self, params: Params = None, device: int = 0):
        L = load()
        self.params = params.copy() if params is not None else default_params()
        h = ctypes.c_void_p()
        _check(L.qpb_create(ctypes.byref(self.params), int(device), ctypes.byref(h)), "qpb_create")
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            load().qpb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(load().qpb_launch_count(self._h))

    # -- device-resident packed records (torch uint8 tensors or raw device pointers) ---------------
    def control_packed(self, d_states, d_out, n, stream=None):
        _check(load().qpb_control_batch_packed(self._h, int(n), _ptr(d_states), _ptr(d_out), stream), "qpb_control_batch_packed")

    # -- device-resident argument-per-array form ----------------------------------------------------
    def control_split(self, n, Rwb, Rwb_d, x, xdot, w, x_d, xdot_d, w_d, feet, contact, q, grf, tau=None, status=None,
                      stream=None):
        args = [_ptr(a) for a in (Rwb, Rwb_d, x, xdot, w, x_d, xdot_d, w_d, feet, contact, q, grf, tau, status)]
        _check(load().qpb_control_batch(self._h, int(n), *args, stream), "qpb_control_batch")

    # -- host buffers: the call a reference-side binding makes --------------------------------------
    def control_host(self, states: np.ndarray, out: np.ndarray = None):
        states = np.ascontiguousarray(states)
        assert states.dtype == STATE_DTYPE
        if out is None:
            out = np.empty(states.shape[0], dtype=OUT_DTYPE)
        assert out.dtype == OUT_DTYPE and out.shape[0] == states.shape[0] and out.flags.c_contiguous
        _check(load().qpb_control_batch_host(self._h, states.shape[0], states.ctypes.data, out.ctypes.data), "qpb_control_batch_host")
        return out

    def control_host_async(self, states: np.ndarray, out: np.ndarray):
        """Queue one batch (pinned buffers, see PinnedBuffer) and return; ``out`` is complete after host_sync()."""
        assert states.dtype == STATE_DTYPE and states.flags.c_contiguous
        assert out.dtype == OUT_DTYPE and out.shape[0] == states.shape[0] and out.flags.c_contiguous
        _check(load().qpb_control_batch_host_async(self._h, states.shape[0], states.ctypes.data, out.ctypes.data),
               "qpb_control_batch_host_async")

    def control_wire_host(self, states: np.ndarray, out: np.ndarray = None):
        """Host buffers in the wire format (488 B up, 200 B down per robot; records.to_wire / from_wire convert)."""
        states = np.ascontiguousarray(states)
        assert states.dtype == WIRE_STATE_DTYPE
        if out is None:
            out = np.empty(states.shape[0], dtype=WIRE_OUT_DTYPE)
        assert out.dtype == WIRE_OUT_DTYPE and out.shape[0] == states.shape[0] and out.flags.c_contiguous
        _check(load().qpb_control_batch_wire_host(self._h, states.shape[0], states.ctypes.data, out.ctypes.data),
               "qpb_control_batch_wire_host")
        return out

    def control_wire_host_async(self, states: np.ndarray, out: np.ndarray):
        assert states.dtype == WIRE_STATE_DTYPE and states.flags.c_contiguous
        assert out.dtype == WIRE_OUT_DTYPE and out.shape[0] == states.shape[0] and out.flags.c_contiguous
        _check(load().qpb_control_batch_wire_host_async(self._h, states.shape[0], states.ctypes.data, out.ctypes.data),
               "qpb_control_batch_wire_host_async")

    def host_sync(self):
        _check(load().qpb_host_sync(self._h), "qpb_host_sync")

    def set_warm_batches(self, on: bool = True):
        """Device-resident calls carry warm-start words (state.pad[0:4] = last tick's out.pad[0:4]): one-launch kernel."""
        _check(load().qpb_set_warm_batches(self._h, 1 if on else 0), "qpb_set_warm_batches")

    # -- whole control tick: balance QP for stance legs + joint PD for swing legs (commander_node.cpp:482-533) --
    def set_joint_gains(self, gains: JointGains):
        _check(load().qpb_set_joint_gains(self._h, ctypes.byref(gains)), "qpb_set_joint_gains")

    def tick_packed(self, d_states, d_swing, d_out, n, stream=None):
        _check(load().qpb_tick_batch_packed(self._h, int(n), _ptr(d_states), _ptr(d_swing), _ptr(d_out), stream), "qpb_tick_batch_packed")

    def tick_host(self, states: np.ndarray, swing: np.ndarray, out: np.ndarray = None):
        states, swing = np.ascontiguousarray(states), np.ascontiguousarray(swing)
        assert states.dtype == STATE_DTYPE and swing.dtype == SWING_DTYPE and len(states) == len(swing)
        if out is None:
            out = np.empty(states.shape[0], dtype=OUT_DTYPE)
        assert out.dtype == OUT_DTYPE and out.shape[0] == states.shape[0] and out.flags.c_contiguous
        _check(load().qpb_tick_batch_host(self._h, states.shape[0], states.ctypes.data, swing.ctypes.data, out.ctypes.data), "qpb_tick_batch_host")
        return out

    # -- the caller code either side of the tick: planner + swing trajectory, message adapters (device pointers) --
    def set_plan_params(self, pp: PlanParams):
        _check(load().qpb_set_plan_params(self._h, ctypes.byref(pp)), "qpb_set_plan_params")

    def plan(self, d_states, d_plan, d_swing, n, stream=None):
        _check(load().qpb_plan_batch(self._h, int(n), _ptr(d_states), _ptr(d_plan), _ptr(d_swing), stream), "qpb_plan_batch")

    def adapt_inputs(self, d_com, d_joints, d_states, d_swing, n, stream=None):
        _check(load().qpb_adapt_inputs_batch(self._h, int(n), _ptr(d_com), _ptr(d_joints), _ptr(d_states), _ptr(d_swing), stream),
               "qpb_adapt_inputs_batch")

    def torque_cmd(self, d_states, d_out, d_cmd, n, stream=None):
        _check(load().qpb_torque_cmd_batch(self._h, int(n), _ptr(d_states), _ptr(d_out), _ptr(d_cmd), stream), "qpb_torque_cmd_batch")

    def jt(self, n, q, grf, contact, tau, stream=None):
        _check(load().qpb_jt_batch(self._h, int(n), _ptr(q), _ptr(grf), _ptr(contact), _ptr(tau), stream), "qpb_jt_batch")

    def jt_host(self, q, grf, contact=None):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, 12)
        grf = np.ascontiguousarray(grf, dtype=np.float64).reshape(-1, 12)
        if contact is not None:
            contact = np.ascontiguousarray(contact, dtype=np.uint8).reshape(-1, 4)
        tau = np.empty_like(q)
        _check(load().qpb_jt_batch_host(self._h, q.shape[0], _ptr(q), _ptr(grf), _ptr(contact), _ptr(tau)), "qpb_jt_batch_host")
        return tau

    def fk_host(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64).reshape(-1, 12)
        feet = np.empty_like(q)
        _check(load().qpb_fk_batch_host(self._h, q.shape[0], _ptr(q), _ptr(feet)), "qpb_fk_batch_host")
        return feet

    def fk(self, n, q, feet, stream=None):
        _check(load().qpb_fk_batch(self._h, int(n), _ptr(q), _ptr(feet), stream), "qpb_fk_batch")


def default_mpc_params():
    p = MpcParams()
    _check(load().qpb_mpc_default_params(ctypes.byref(p)), "qpb_mpc_default_params")
    return p


class MpcSolver:
    """Owns one ``qpb_mpc_handle``: the batched 10-step convex-MPC QP (BASELINE config 4)."""

    def __init__(self, params: MpcParams = None, device: int = 0):
        L = load()
        self.params = params.copy() if params is not None else default_mpc_params()
        h = ctypes.c_void_p()
        _check(L.qpb_mpc_create(ctypes.byref(self.params), int(device), ctypes.byref(h)), "qpb_mpc_create")
        self._h = h
        self.device = int(device)

    def close(self):
        if getattr(self, "_h", None):
            load().qpb_mpc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        return int(load().qpb_mpc_launch_count(self._h))

    def solve_packed(self, d_recs, d_out, n, stream=None):
        _check(load().qpb_mpc_batch_packed(self._h, int(n), _ptr(d_recs), _ptr(d_out), stream), "qpb_mpc_batch_packed")

    def solve_host(self, recs: np.ndarray, out: np.ndarray = None):
        recs = np.ascontiguousarray(recs)
        assert recs.dtype == MPC_REC_DTYPE
        if out is None:
            out = np.empty(recs.shape[0], dtype=MPC_OUT_DTYPE)
        assert out.dtype == MPC_OUT_DTYPE and out.shape[0] == recs.shape[0] and out.flags.c_contiguous
        _check(load().qpb_mpc_batch_host(self._h, recs.shape[0], recs.ctypes.data, out.ctypes.data), "qpb_mpc_batch_host")
        return out
