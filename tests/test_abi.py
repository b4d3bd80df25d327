"""CPU tests of the drop-in boundary: the shared library loads, exports every symbol the header
declares, has the declared struct layouts, validates parameters, and fails loudly without a GPU."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from quadruped_control_b200 import OUT_DTYPE, STATE_DTYPE, Params, default_params, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "qpb200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qpb_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    L = lib.load()
    names = _declared_functions()
    assert len(names) >= 15
    for name in names:
        assert hasattr(L, name), f"{name} declared in qpb200.h but not exported"
    assert set(names) == set(lib.EXPORTS)
    assert L.qpb_version() == 200


def test_struct_layouts_match_header(built, tmp_path):
    """Compile the header as C and compare sizeof/offsetof with the numpy/ctypes mirrors."""
    prog = tmp_path / "layout.c"
    prog.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "qpb200.h"\n'
        "int main(void){printf(\"%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(qpb_params), sizeof(qpb_state_rec),"
        " sizeof(qpb_out_rec), offsetof(qpb_state_rec, feet), offsetof(qpb_state_rec, q), offsetof(qpb_state_rec, contact),"
        " offsetof(qpb_out_rec, tau), offsetof(qpb_out_rec, status), offsetof(qpb_params, max_iter));return 0;}\n"
    )
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert got[0] == ctypes.sizeof(Params)
    assert got[1] == STATE_DTYPE.itemsize == 512 and got[2] == OUT_DTYPE.itemsize == 256
    assert got[3] == STATE_DTYPE.fields["feet"][1] and got[4] == STATE_DTYPE.fields["q"][1]
    assert got[5] == STATE_DTYPE.fields["contact"][1] == 480
    assert got[6] == OUT_DTYPE.fields["tau"][1] and got[7] == OUT_DTYPE.fields["status"][1] == 192
    assert got[8] == Params.max_iter.offset


def test_default_params_agree_with_python_mirror(built):
    assert bytes(lib.default_params()) == bytes(default_params(0.8))


def _create(p):
    h = ctypes.c_void_p()
    rc = lib.load().qpb_create(ctypes.byref(p), 0, ctypes.byref(h))
    if rc == 0:
        lib.load().qpb_destroy(h)
    return rc, lib.load().qpb_last_error().decode()


@pytest.mark.parametrize(
    "mutate, text",
    [
        (lambda p: setattr(p, "mu", 0.0), "mu"),
        (lambda p: setattr(p, "mu", float("nan")), "non-finite"),
        (lambda p: setattr(p, "fzmin", 200.0), "fzmin"),
        (lambda p: (setattr(p, "fzmin", -5.0), setattr(p, "fzmax", -1.0)), "fzmax"),
        (lambda p: setattr(p, "fzmax", 1e7), "1e6"),
        (lambda p: setattr(p, "max_iter", 0), "max_iter"),
        (lambda p: p.W.__setitem__(0, -1.0), "W"),
        (lambda p: p.S.__setitem__(1, 0.5), "S"),
    ],
)
def test_create_rejects_bad_parameters_before_touching_cuda(built, mutate, text):
    p = default_params(0.6)
    mutate(p)
    rc, msg = _create(p)
    assert rc == -2 and text in msg, (rc, msg)


def test_null_arguments(built):
    L = lib.load()
    assert L.qpb_create(None, 0, None) == -1
    assert L.qpb_default_params(None) == -1
    assert L.qpb_control_batch_host(None, 1, None, None) == -1
    assert L.qpb_destroy(None) == 0
    assert L.qpb_launch_count(None) == 0


def test_no_cpu_fallback(built):
    """Without a CUDA device the constructor must fail loudly (QPB_ERR_CUDA), never compute on the CPU."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the -m gpu tests")
    with pytest.raises(lib.QpbError) as ei:
        lib.BalanceSolver(default_params(0.6))
    assert "(-3)" in str(ei.value)


def test_product_path_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "quadruped_control_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "qpb_oracle" not in text and "from oracle" not in text, f
    assert "oracle" not in open(HEADER).read()


def test_cpp_shim_compiles_against_the_abi(built):
    import __graft_entry__ as g

    exe = g.build_cpp_shim_check()
    assert exe and os.path.exists(exe)


def test_multi_shard_ranges_match_the_torchrun_sharding(built):
    """The single-process multi-GPU calls (qpb_multi_*) cut the batch exactly as sharding.shard_range does."""
    from quadruped_control_b200.sharding import shard_range

    for n in (0, 1, 7, 8, 1001, 65536, 8388608 + 3):
        for g in (1, 2, 3, 8):
            got = [lib.multi_shard_range(n, r, g) for r in range(g)]
            assert got == [shard_range(n, r, g) for r in range(g)]
            assert got[0][0] == 0 and got[-1][1] == n and all(got[r][1] == got[r + 1][0] for r in range(g - 1))
    L = lib.load()
    lo, hi = ctypes.c_int64(), ctypes.c_int64()
    assert L.qpb_multi_shard_range(10, 2, 2, ctypes.byref(lo), ctypes.byref(hi)) == -1
    assert L.qpb_multi_shard_range(-1, 0, 2, ctypes.byref(lo), ctypes.byref(hi)) == -1


def test_multi_create_fails_loudly_without_a_gpu(built):
    import torch

    L = lib.load()
    assert L.qpb_multi_create(None, None, 0, None) == -1
    assert L.qpb_multi_destroy(None) == 0 and L.qpb_multi_num_shards(None) == 0 and L.qpb_multi_launch_count(None) == 0
    assert L.qpb_multi_control_batch_host(None, 1, None, None) == -1
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the -m gpu tests")
    with pytest.raises(lib.QpbError) as ei:
        lib.MultiBalanceSolver(default_params(0.6))
    assert "(-3)" in str(ei.value)
    with pytest.raises(lib.QpbError):
        lib.MultiBalanceSolver(default_params(0.6), devices=[0, 0])
