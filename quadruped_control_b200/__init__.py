"""qpb200: Blackwell-native batched balance-controller QP solver."""
from .records import (  # noqa: F401
    ALGO_BYTES_PER_QP,
    LEG_NAMES,
    OUT_DTYPE,
    QPB_BAD_INPUT,
    QPB_MAX_ITER,
    QPB_OK,
    STATE_DTYPE,
    SWING_DTYPE,
    WIRE_OUT_DTYPE,
    WIRE_STATE_DTYPE,
    from_wire,
    to_wire,
    JointGains,
    default_joint_gains,
    Params,
    default_params,
)
