"""Host sanitizers for the CPU oracle: the plain-C restatement is rebuilt with
AddressSanitizer + UndefinedBehaviorSanitizer and run over balance, whole-tick and MPC records -- including non-finite
inputs, every contact mask and degenerate bounds -- and its results must equal the optimised build the other tests use."""
import ctypes
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle
from quadruped_control_b200 import OUT_DTYPE, default_params, states
from quadruped_control_b200.records import MPC_OUT_DTYPE, default_mpc_params

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def san_binary(tmp_path_factory):
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    exe = str(tmp_path_factory.mktemp("san") / "oracle_san")
    srcs = [os.path.join(HERE, "sanitize", "oracle_san_main.c")] + [os.path.join(ROOT, "oracle", f) for f in ("qpb_oracle.c", "mpc_oracle.c", "plan_oracle.c")]
    cmd = ["gcc", "-O1", "-g", "-std=gnu99", "-fno-omit-frame-pointer", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
           "-o", exe] + srcs + ["-lm", "-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0 and "sanitize" in res.stderr:
        pytest.skip("this gcc has no sanitizer runtime: " + res.stderr[-300:])
    assert res.returncode == 0, res.stderr[-2000:]
    return exe


def test_oracle_is_clean_under_asan_and_ubsan(san_binary, tmp_path):
    p = default_params(0.6)
    p.fzmin = 0.0  # pyramid apex: the degenerate end of the active-set solver
    S = states.generate_states(600, 424242, profile="stress", masks="mixed")
    S["contact"][:160] = (np.arange(160)[:, None] % 16 >> np.arange(4)) & 1
    S["xdot"][7, 1] = np.nan
    S["Rwb"][11, 4] = np.inf
    sw = states.generate_swing(S, 5, p)
    mp = default_mpc_params(0.6)
    R = states.generate_mpc(24, 20260104)
    inp, outp = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as f:
        f.write(np.array([len(S), len(R)], dtype=np.int64).tobytes())
        f.write(bytes(p))
        f.write(S.tobytes())
        f.write(sw.tobytes())
        f.write(bytes(mp))
        f.write(R.tobytes())
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=1:abort_on_error=0", UBSAN_OPTIONS="print_stacktrace=1")
    res = subprocess.run([san_binary, inp, outp], capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, (res.returncode, res.stderr[-3000:])
    assert "runtime error" not in res.stderr and "AddressSanitizer" not in res.stderr, res.stderr[-3000:]
    raw = open(outp, "rb").read()
    nb = len(S) * OUT_DTYPE.itemsize
    out = np.frombuffer(raw[:nb], dtype=OUT_DTYPE)
    tick = np.frombuffer(raw[nb:2 * nb], dtype=OUT_DTYPE)
    mout = np.frombuffer(raw[2 * nb:], dtype=MPC_OUT_DTYPE)
    ref = oracle.control_batch(p, S, 2)
    assert np.array_equal(out["status"], ref["status"]) and list(np.nonzero(out["status"])[0]) == [7, 11]
    assert np.allclose(out["grf_body"], ref["grf_body"], rtol=0, atol=1e-9) and np.allclose(out["tau"], ref["tau"], rtol=0, atol=1e-9)
    rt = oracle.tick_batch(p, oracle.default_joint_gains(), S, sw, 2)
    assert np.array_equal(tick["status"], rt["status"]) and np.allclose(tick["tau"], rt["tau"], rtol=0, atol=1e-8)
    rm = oracle.mpc_batch(mp, R, 2)
    assert np.array_equal(mout["status"], rm["status"]) and np.allclose(mout["U"], rm["U"], rtol=0, atol=1e-8)
