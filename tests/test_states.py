import numpy as np

from quadruped_control_b200 import STATE_DTYPE, default_params, states
from quadruped_control_b200.sharding import shard_range


def test_generator_is_index_addressable():
    full = states.generate_states(1000, 20260103, masks="mixed")
    for lo, n in ((0, 10), (123, 456), (999, 1)):
        part = states.generate_states(n, 20260103, lo=lo, masks="mixed")
        assert part.tobytes() == full[lo:lo + n].tobytes()
    assert states.generate_states(8, 1).tobytes() != states.generate_states(8, 2).tobytes()


def test_shards_tile_the_stream():
    full = states.generate_states(1003, 5)
    parts = []
    for r in range(8):
        lo, hi = shard_range(1003, r, 8)
        parts.append(states.generate_states(hi - lo, 5, lo=lo))
    assert np.concatenate(parts).tobytes() == full.tobytes()
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]


def test_distributions_match_survey_8d():
    S = states.generate_states(20000, 20260102)
    assert S.dtype == STATE_DTYPE and (S["contact"] == 1).all()
    R = S["Rwb"].reshape(-1, 3, 3)
    assert np.allclose(R @ R.transpose(0, 2, 1), np.eye(3), atol=1e-12)
    assert np.allclose(np.linalg.det(R), 1.0, atol=1e-12)
    Rd = S["Rwb_d"].reshape(-1, 3, 3)
    ang = np.arccos(np.clip((np.trace(Rd @ R.transpose(0, 2, 1), axis1=1, axis2=2) - 1) / 2, -1, 1))
    assert ang.max() <= 0.05 + 1e-9 and ang.mean() > 0.02
    assert np.abs(S["x"] - [0, 0, 0.26]).max() <= 0.05 and np.abs(S["x_d"] - S["x"]).max() <= 0.05 + 1e-12
    assert abs(S["xdot"].std() - 0.3) < 0.01 and abs(S["w"].std() - 0.5) < 0.02
    assert abs((S["w_d"] - S["w"]).std() - 0.2) < 0.01
    assert np.abs(S["q"] - states.STANCE_Q).max() <= 0.3
    assert np.allclose(S["feet"], states.forward_kinematics(S["q"], default_params()))
    M = states.generate_states(22000, 20260103, masks="mixed")
    codes = (M["contact"] * [1, 2, 4, 8]).sum(axis=1)
    counts = np.bincount(codes, minlength=16)
    assert (counts[[m for m in range(16) if bin(m).count("1") < 2]] == 0).all()
    assert counts[states.MIXED_MASKS].min() > 1700  # uniform over the 11 masks


def test_wire_records_round_trip():
    """to_wire / from_wire (qpb_wire_state 488 B, qpb_wire_out 200 B): the device records without their padding."""
    from quadruped_control_b200 import OUT_DTYPE, WIRE_OUT_DTYPE, WIRE_STATE_DTYPE, from_wire, to_wire

    S = states.generate_states(37, 5, masks="mixed")
    S["pad"][:, :4] = np.arange(37 * 4, dtype=np.uint8).reshape(37, 4)
    W = to_wire(S)
    assert W.dtype == WIRE_STATE_DTYPE and W.itemsize == 488
    # the wire record IS the first 488 bytes of the device record
    assert W.tobytes() == b"".join(S[i].tobytes()[:488] for i in range(37))
    o = np.zeros(5, dtype=WIRE_OUT_DTYPE)
    o["grf_body"] = np.arange(60).reshape(5, 12)
    o["status"], o["iters"], o["wset"] = [0, 1, 2, 0, 0], [3, 200, 0, 7, 32767], [0x80000001, 0x80ffffff, 0, 5, 6]
    back = from_wire(o)
    assert back.dtype == OUT_DTYPE and np.array_equal(back["iters"], o["iters"]) and np.array_equal(back["status"], o["status"])
    assert np.array_equal(back["pad"][:, :4].copy().view("<u4")[:, 0], o["wset"]) and not back["pad"][:, 4:].any()
