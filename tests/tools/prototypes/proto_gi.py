"""numpy prototype of the operator-form dual active-set that the CUDA kernel implements."""
import numpy as np, sys
sys.path.insert(0,'/root/repo')
import oracle
from quadruped_control_b200 import default_params, states

def cons_table(mu, fzmin, fzmax):
    # per type: (ia_off, ca, ib_off, cb, bound)
    return [(0,-1.0,2,mu,0.0),(1,-1.0,2,mu,0.0),(1,1.0,2,mu,0.0),(0,1.0,2,mu,0.0),(2,1.0,2,0.0,fzmin),(2,-1.0,2,0.0,-fzmax)]

def solve(Q, c, contact, mu, fzmin, fzmax, max_iter=200, eps_dep=1e-10, tol=1e-9, stats=None):
    Q = Q.copy(); c=c.copy()
    for leg in range(4):
        if not contact[leg]:
            for i in range(3*leg,3*leg+3):
                Q[i,:]=0; Q[:,i]=0; Q[i,i]=1.0; c[i]=0.0
    H = np.linalg.inv(Q)
    x = -H@c
    tab = cons_table(mu,fzmin,fzmax)
    Nrm = np.zeros((24,12)); b=np.zeros(24); en=np.zeros(24,bool)
    for j in range(24):
        leg,t=divmod(j,6); ia,ca,ib,cb,bd = tab[t]
        Nrm[j,3*leg+ia]+=ca; Nrm[j,3*leg+ib]+=cb; b[j]=bd; en[j]=bool(contact[leg])
    hnn = np.einsum('ji,ik,jk->j',Nrm,H,Nrm)
    P=H.copy(); Ns=np.zeros((12,12)); slot=-np.ones(12,int); u=np.zeros(12); active=np.zeros(24,bool)
    p=-1; up=0.0; it=0; status=0
    while True:
        if p<0:
            s = Nrm@x-b
            s[~en]=np.inf; s[active]=np.inf
            p=int(np.argmin(s))
            if s[p] >= -tol*(1+abs(b[p])): break
            up=0.0
        if it>=max_iter: status=1; break
        it+=1
        n=Nrm[p]
        z=P@n; r=Ns@n
        zeta=n@z
        dep = zeta <= eps_dep*hnn[p]
        sp = n@x-b[p]
        act = slot>=0
        t1=np.inf; k=-1
        for kk in range(12):
            if act[kk] and r[kk]>0:
                tt=u[kk]/r[kk]
                if tt<t1: t1=tt;k=kk
        t2 = np.inf if dep else max(0.0,-sp/zeta)
        t=min(t1,t2)
        if t==np.inf: status=2; break
        if not dep: x=x+t*z
        u[act]-=t*r[act]; up+=t
        if t2<=t1:
            q=int(np.argmin(slot>=0))
            P-=np.outer(z,z)/zeta
            Ns[act]-=np.outer(r[act]/zeta,z)
            Ns[q]=z/zeta; slot[q]=p; u[q]=up; active[p]=True; p=-1
        else:
            nu=Ns[k].copy(); Qnu=Q@nu; delta=nu@Qnu; gam=Ns@Qnu
            P+=np.outer(nu,nu)/delta
            Ns-=np.outer(gam/delta,nu)
            Ns[k]=0; active[slot[k]]=False; slot[k]=-1; u[k]=0
            if stats is not None: stats['drops']=stats.get('drops',0)+1
    return status,x,it

if __name__=="__main__":
    p6=default_params(0.6)
    for masks,seed in (("all4",20260102),("mixed",20260103)):
        for prof in ("default","light","stress"):
            S=states.generate_states(3000,seed,profile=prof,masks=masks)
            worst=0; its=[]; oits=[]; stats={}
            for i in range(len(S)):
                qp=oracle.assemble(p6,S[i:i+1])
                st,xo,lam,ito=oracle.qp_solve(qp['Q'],qp['c'],qp['C'],qp['lb'],qp['ub'])
                st2,x,it=solve(qp['Q'],qp['c'],S['contact'][i],0.6,10.0,120.0,stats=stats)
                assert st2==0,(i,st2)
                err=np.abs(x-xo).max()/max(np.abs(xo).max(),1.0)
                worst=max(worst,err); its.append(it); oits.append(ito)
            print(masks,prof,'worst rel err',worst,'mean it',np.mean(its),'max',max(its),'oracle mean it',np.mean(oits),'drops/QP',stats.get('drops',0)/len(S))
