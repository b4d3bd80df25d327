"""Lock-step cost of solving two QPs per warp, and what smarter pairing could recover."""
import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests/tools/prototypes')
import oracle
from quadruped_control_b200 import default_params, states
from proto_rules import solve
p6=default_params(0.6)
n=4096
for name,masks,seed,profile in (("cfg2","all4",20260102,"default"),("cfg3","mixed",20260103,"default"),("tick-light","mixed",20260103,"light")):
    S=states.generate_states(n,seed,masks=masks,profile=profile)
    its=np.zeros(n,int); viol0=np.zeros(n,int)
    for i in range(n):
        qp=oracle.assemble(p6,S[i:i+1])
        st,x,it,a,d=solve(qp['Q'],qp['c'],S['contact'][i],0.6,10.0,120.0,'raw')
        its[i]=it
    nst=S['contact'].astype(bool).sum(1)
    lock=np.maximum(its[0::2],its[1::2])
    print(name,'mean iters %.2f; lockstep per QP %.2f (waste %.0f%%)'%(its.mean(), lock.mean(), 100*(lock.mean()/max(its.mean(),1e-9)-1)))
    for W in (8,16,32):
        tot=0; tot_or=0
        for w0 in range(0,n,W):
            idx=np.arange(w0,w0+W)
            o=idx[np.argsort(nst[idx],kind='stable')]
            tot+=np.maximum(its[o[0::2]],its[o[1::2]]).sum()
            o2=idx[np.argsort(its[idx],kind='stable')]
            tot_or+=np.maximum(its[o2[0::2]],its[o2[1::2]]).sum()
        print('   window %d: pair by #stance %.2f, oracle pairing %.2f'%(W, tot/(n/2), tot_or/(n/2)))
    for k in (2,3,4):
        m=nst==k
        if m.any(): print('   #stance',k,'mean iters %.2f sd %.2f'%(its[m].mean(), its[m].std()))
