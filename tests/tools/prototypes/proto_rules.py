"""Iteration counts of the whitened dual active set under different entering-row rules."""
import numpy as np, sys
sys.path.insert(0,"/root/repo"); sys.path.insert(0,"/root/repo/tests/tools/prototypes")
import oracle
from quadruped_control_b200 import default_params, states
from proto_gi import cons_table

def solve(Q, c, contact, mu, fzmin, fzmax, rule, max_iter=200, eps_dep=1e-13, tol=1e-9):
    Q = Q.copy(); c=c.copy()
    for leg in range(4):
        if not contact[leg]:
            for i in range(3*leg,3*leg+3):
                Q[i,:]=0; Q[:,i]=0; Q[i,i]=1.0; c[i]=0.0
    L = np.linalg.cholesky(Q); J0 = np.linalg.inv(L).T
    y = -(J0.T@c); x = J0@y
    tab = cons_table(mu,fzmin,fzmax)
    Nrm = np.zeros((24,12)); b=np.zeros(24); en=np.zeros(24,bool)
    for j in range(24):
        leg,t=divmod(j,6); ia,ca,ib,cb,bd = tab[t]
        Nrm[j,3*leg+ia]+=ca; Nrm[j,3*leg+ib]+=cb; b[j]=bd; en[j]=bool(contact[leg])
    Nt = Nrm@J0   # rows: whitened normals
    nn = (Nt*Nt).sum(1)
    P=np.eye(12); Ns=np.zeros((12,12)); slot=-np.ones(12,int); u=np.zeros(12); active=np.zeros(24,bool)
    p=-1; up=0.0; it=0; adds=0; drops=0
    while True:
        if p<0:
            s = Nrm@x-b
            viol = en & ~active & (s < -tol*(1+np.abs(b)))
            if not viol.any(): break
            if rule=='raw': score = s
            elif rule=='norm': score = s/np.sqrt(nn)
            elif rule=='proj':
                zz = np.einsum('ji,ik,jk->j',Nt,P,Nt); score = s/np.sqrt(np.maximum(zz,1e-300))
            elif rule=='least': score = -s
            elif rule=='first': score = np.arange(24.0)
            score = np.where(viol, score, np.inf)
            p=int(np.argmin(score)); up=0.0
        if it>=max_iter: return 1,x,it,adds,drops
        it+=1
        nt = Nt[p]; z=P@nt; r=Ns@nt; zeta=nt@z
        dep = zeta <= eps_dep*nn[p]
        sp = Nrm[p]@x-b[p]
        act = slot>=0
        t1=np.inf; k=-1
        for kk in range(12):
            if act[kk] and r[kk]>0:
                tt=u[kk]/r[kk]
                if tt<t1: t1=tt;k=kk
        t2 = np.inf if dep else max(0.0,-sp/zeta)
        t=min(t1,t2)
        if t==np.inf: return 2,x,it,adds,drops
        if not dep: x = x + t*(J0@z)
        u[act]-=t*r[act]; up+=t
        if t2<=t1:
            q=int(np.argmin(slot>=0))
            P-=np.outer(z,z)/zeta; Ns[act]-=np.outer(r[act]/zeta,z)
            Ns[q]=z/zeta; slot[q]=p; u[q]=up; active[p]=True; p=-1; adds+=1
        else:
            nu=Ns[k].copy(); delta=nu@nu; gam=Ns@nu
            P+=np.outer(nu,nu)/delta; Ns-=np.outer(gam/delta,nu)
            Ns[k]=0; active[slot[k]]=False; slot[k]=-1; u[k]=0; drops+=1
    return 0,x,it,adds,drops

if __name__=="__main__":
    p6=default_params(0.6)
    n=int(sys.argv[1]) if len(sys.argv)>1 else 1500
    for masks,seed in (("all4",20260102),("mixed",20260103)):
        S=states.generate_states(n,seed,masks=masks)
        qps=[oracle.assemble(p6,S[i:i+1]) for i in range(n)]
        for rule in ('raw','norm','proj','least','first'):
            its=[];ad=[];dr=[]; worst=0
            for i in range(n):
                st,x,it,a,d=solve(qps[i]['Q'],qps[i]['c'],S['contact'][i],0.6,10.0,120.0,rule)
                assert st==0
                its.append(it); ad.append(a); dr.append(d)
            print(masks, rule, 'iters mean %.2f max %d  adds %.2f drops %.2f  sd %.2f'%(np.mean(its),max(its),np.mean(ad),np.mean(dr),np.std(its)))
