import numpy as np, sys
sys.path.insert(0,"/root/repo"); sys.path.insert(0,"/root/repo/tests/tools/prototypes")
import oracle
from quadruped_control_b200 import default_params, states
from proto_gi import cons_table
p6=default_params(0.6)
n=800
for masks,seed in (("all4",20260102),("mixed",20260103)):
    S=states.generate_states(n,seed,masks=masks)
    stats=[]
    for i in range(n):
        qp=oracle.assemble(p6,S[i:i+1])
        st,xo,lam,ito=oracle.qp_solve(qp['Q'],qp['c'],qp['C'],qp['lb'],qp['ub'])
        Q=qp['Q']; c=qp['c']; contact=S['contact'][i]
        # final active set in 24-row numbering of proto (leg*6+type): types 0:-fx+mu fz,1:-fy+mu fz,2:fy+mu fz,3:fx+mu fz,4:fz>=min,5:fz<=max
        tab=cons_table(0.6,10.0,120.0)
        Nrm=np.zeros((24,12)); b=np.zeros(24); en=np.zeros(24,bool)
        for j in range(24):
            leg,t=divmod(j,6); ia,ca,ib,cb,bd=tab[t]
            Nrm[j,3*leg+ia]+=ca; Nrm[j,3*leg+ib]+=cb; b[j]=bd; en[j]=bool(contact[leg])
        s_opt=Nrm@xo-b
        act_opt = en & (np.abs(s_opt)<1e-7)
        # unconstrained (stance-only) solution
        idx=[k for k in range(12) if contact[k//3]]
        x0=np.zeros(12); x0[idx]=-np.linalg.solve(Q[np.ix_(idx,idx)],c[idx])
        s0=Nrm@x0-b
        viol0 = en & (s0< -1e-9)
        stats.append((act_opt.sum(), viol0.sum(), (viol0&act_opt).sum(), (viol0&~act_opt).sum(), (act_opt&~viol0).sum()))
    st=np.array(stats)
    print(masks,'|A*| %.2f  |V0| %.2f  V0&A* %.2f  V0 not in A* %.2f  A* not in V0 %.2f'%tuple(st.mean(0)))
    print('   exact match frac', np.mean((st[:,3]==0)&(st[:,4]==0)), ' V0 subset of A* frac', np.mean(st[:,3]==0))
