"""compute-sanitizer over every kernel family: memcheck (out-of-bounds / misaligned
accesses) and racecheck (shared-memory hazards: the loop kernel's per-QP side blocks and staging buffer, the half-warp
kernels' broadcast buffers, the MPC kernel's pivot columns) on tools/sanitize_run.py."""
import os
import shutil
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sanitizer():
    for cand in (shutil.which("compute-sanitizer"), "/usr/local/cuda/bin/compute-sanitizer"):
        if cand and os.path.exists(cand):
            return cand
    pytest.fail("compute-sanitizer not found in this CUDA toolkit")


@pytest.mark.parametrize("tool,size", [("memcheck", "701"), ("racecheck", "257")])
def test_every_kernel_family_is_clean(built, tool, size):
    res = subprocess.run([_sanitizer(), "--tool", tool, "--print-limit", "8", "--error-exitcode", "9", sys.executable,
                          os.path.join(ROOT, "tools", "sanitize_run.py"), size], capture_output=True, text=True, timeout=1500, cwd=ROOT)
    log = os.path.join(ROOT, "gpurun_out", f"sanitizer_{tool}.log")  # kept for the report when the box writes gpurun_out/
    try:
        os.makedirs(os.path.dirname(log), exist_ok=True)
        open(log, "w").write(res.stdout + "\n==== stderr ====\n" + res.stderr)
    except OSError:
        pass
    lines = [l for l in (res.stdout + res.stderr).splitlines() if "=========" in l]
    tail = "\n".join(lines[:40]) + "\n...\n" + (res.stdout + res.stderr)[-1500:]
    assert res.returncode == 0, tail
    assert "sanitize_run ok" in res.stdout, tail
    assert "ERROR SUMMARY: 0 errors" in res.stdout + res.stderr or "RACECHECK SUMMARY: 0 hazards" in res.stdout + res.stderr, tail
