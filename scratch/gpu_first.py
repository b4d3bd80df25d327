import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
import torch
import oracle
from quadruped_control_b200 import lib, states, default_params, STATE_DTYPE, OUT_DTYPE

def run(params, S, tag):
    sol = lib.BalanceSolver(params)
    out = sol.control_host(S)
    ref = oracle.control_batch(params, S, 8)
    den = np.maximum(np.abs(ref['grf_body']).max(axis=1), 1.0)
    ef = np.abs(out['grf_body']-ref['grf_body']).max(axis=1)/den
    dent = np.maximum(np.abs(ref['tau']).max(axis=1), 1.0)
    et = np.abs(out['tau']-ref['tau']).max(axis=1)/dent
    print(tag, 'n', len(S), 'status', np.bincount(out['status']), 'ref status', np.bincount(ref['status']),
          'max rel grf', ef.max(), 'tau', et.max(), 'iters mean', out['iters'].mean(), 'max', out['iters'].max(), 'ref iters', ref['iters'].mean())
    bad = np.argsort(-ef)[:3]
    for b in bad:
        if ef[b] > 1e-6:
            print('  worst', b, ef[b], out['status'][b], out['iters'][b]); print('   gpu', out['grf_body'][b]); print('   ref', ref['grf_body'][b])
    return sol

p8 = default_params(0.8)
S1 = states.stance_state(p8)
sol = run(p8, S1, 'cfg1')
p6 = default_params(0.6)
for prof in ('default','light','stress'):
    run(p6, states.generate_states(4096, 20260102, profile=prof), 'all4/'+prof)
    run(p6, states.generate_states(4096, 20260103, profile=prof, masks='mixed'), 'mixed/'+prof)
# timing
S = states.generate_states(65536, 20260102)
sol = lib.BalanceSolver(p6)
d_in = torch.from_numpy(S.view(np.uint8).reshape(-1)).cuda()
d_out = torch.empty(len(S)*256, dtype=torch.uint8, device='cuda')
for _ in range(3): sol.control_packed(d_in, d_out, len(S))
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): sol.control_packed(d_in, d_out, len(S))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/10
print('kernel ms', ms, 'QP/s', len(S)/ms*1e3)
