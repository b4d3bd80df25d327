import numpy as np, sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/scratch')
import oracle
from quadruped_control_b200 import default_params, states
import proto_gi
p6=default_params(0.6)
S=states.generate_states(3000,20260102)
errs=[]
for i in range(len(S)):
    qp=oracle.assemble(p6,S[i:i+1])
    st,xo,lam,ito=oracle.qp_solve(qp['Q'],qp['c'],qp['C'],qp['lb'],qp['ub'])
    st2,x,it=proto_gi.solve(qp['Q'],qp['c'],S['contact'][i],0.6,10.0,120.0)
    errs.append(np.abs(x-xo).max()/max(np.abs(xo).max(),1.0))
errs=np.array(errs)
print(np.percentile(errs,[50,90,99,99.9,100]))
i=int(np.argmax(errs)); print('worst',i,errs[i])
qp=oracle.assemble(p6,S[i:i+1])
st,xo,lam,ito=oracle.qp_solve(qp['Q'],qp['c'],qp['C'],qp['lb'],qp['ub'])
st2,x,it=proto_gi.solve(qp['Q'],qp['c'],S['contact'][i],0.6,10.0,120.0)
print(ito,it); print(xo); print(x); print('lam',lam)
H=np.linalg.inv(qp['Q']); print('x0',-H@qp['c'])
print(np.linalg.eigvalsh(qp['Q']))
