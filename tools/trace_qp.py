"""Working-set changes of the longest QPs of BASELINE config 2, one line per change (rows: 3 leg + group, group 0 / 1 / 2 = x / y / z;
A / B = the two sides of a group).  Builds tools/trace_qp.cpp with g++ (the solver core compiles for the host) -- no GPU needed.
  python tools/trace_qp.py [how many QPs to print]"""
import ctypes, os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np
from quadruped_control_b200 import default_params, states

lib = os.path.join(HERE, "..", "scratch", "libtrace_qp.so")
os.makedirs(os.path.dirname(lib), exist_ok=True)
subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", lib, os.path.join(HERE, "trace_qp.cpp")], check=True)
L = ctypes.CDLL(lib)
L.trace_one.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
p = default_params(0.6)
S = states.generate_states(65536, 20260102, masks="all4")
it = np.array([L.trace_one(ctypes.byref(p), S[i:i + 1].ctypes.data, 0) for i in range(len(S))])
print(f"config 2: working-set changes per QP mean {it.mean():.2f}, max {it.max()}; histogram {np.bincount(it).tolist()}")
for i in np.argsort(-it)[:int(sys.argv[1]) if len(sys.argv) > 1 else 2]:
    print(f"QP {i}: {it[i]} changes")
    sys.stdout.flush()
    L.trace_one(ctypes.byref(p), S[i:i + 1].ctypes.data, 1)
