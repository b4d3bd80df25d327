"""The loop pass of BASELINE config 2 as a scheduling problem (no GPU): 1 184 warps of 32 lanes, a worklist of the QPs that
need the loop (their numbers of working-set changes from the host build of the solver core, tools/trace_qp.cpp), a warp
refills its idle lanes when two or more are idle and otherwise does one working-set change for all its lanes in T us
(+ T_refill when it refilled), independently of every other warp.  Prints when the CTAs (two warps) leave, next to the
measured per-CTA timeline (profiles/r02_overlap_ab.txt): the pass is as long as the longest sequence of changes among a warp's
lanes, not as long as its share of the work.  usage: python tools/model_loop_pass.py [T_us] [T_refill_us]"""
import ctypes, os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import numpy as np
from quadruped_control_b200 import default_params, states

T = float(sys.argv[1]) if len(sys.argv) > 1 else 2.4
TRF = float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
lib = os.path.join(HERE, "..", "scratch", "libtrace_qp.so")
os.makedirs(os.path.dirname(lib), exist_ok=True)
subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", lib, os.path.join(HERE, "trace_qp.cpp")], check=True)
L = ctypes.CDLL(lib)
L.trace_counts.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
p = default_params(0.6)
n = 65536
S = states.generate_states(n, 20260102, masks="all4")
su, tot = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32)
L.trace_counts(ctypes.byref(p), S.ctypes.data, n, su.ctypes.data, tot.ctypes.data)
its = (tot - su)[tot > su]  # changes left to the loop pass, in worklist (= record) order
nw = 1184
print(f"config 2: {len(its)} of {n} QPs need the loop; changes in the loop mean {its.mean():.2f}, max {its.max()}; "
      f"{its.sum()} lane-trips on {nw * 32} lanes = {its.sum() / (nw * 32):.1f} per lane = {its.sum() / (nw * 32) * T:.0f} us if they were spread evenly")
rem = np.zeros((nw, 32), dtype=np.int64)
t = np.zeros(nw)
end = np.zeros(nw)
alive = np.ones(nw, bool)
trips = np.zeros(nw, dtype=np.int64)
nxt = 0
while alive.any():
    w = int(np.argmin(np.where(alive, t, 1e18)))  # the warp that is furthest behind takes its next round
    idle = rem[w] == 0
    refilled = False
    if idle.sum() >= 2 or idle.all():
        k = min(int(idle.sum()), len(its) - nxt)
        if k > 0:
            rem[w, np.nonzero(idle)[0][:k]] = its[nxt:nxt + k]
            nxt += k
            refilled = True
    if not rem[w].any():
        alive[w] = False
        end[w] = t[w]
        continue
    rem[w, rem[w] > 0] -= 1
    trips[w] += 1
    t[w] += T + (TRF if refilled else 0.0)
cta = np.maximum(end[0::2], end[1::2])
pct = (0, 10, 25, 50, 75, 90, 99, 100)
print(f"model (T = {T} us per change, {TRF} us per refill round): trips per warp mean {trips.mean():.1f}, max {trips.max()}")
print("  percentile of CTAs     " + "".join(f"{q:>7d}" for q in pct))
print("  model: CTA leaves at   " + "".join(f"{v:7.1f}" for v in np.percentile(cta, pct)))
print("  measured on the B200   " + "".join(f"{v:7.1f}" for v in (28.9, 38.9, 43.0, 46.8, 51.7, 56.6, 65.6, 71.7)))
