"""y-space (whitened) operator-form dual active set: y = L'x, Hessian = I."""
import numpy as np, sys
sys.path.insert(0,"/root/repo"); sys.path.insert(0,"/root/repo/tests/tools/prototypes")
import oracle
from quadruped_control_b200 import default_params, states
from proto_gi import cons_table

def solve(Q, c, contact, mu, fzmin, fzmax, max_iter=200, eps_dep=1e-13, tol=1e-9, stats=None, dtype=np.float64):
    Q = Q.copy(); c=c.copy()
    for leg in range(4):
        if not contact[leg]:
            for i in range(3*leg,3*leg+3):
                Q[i,:]=0; Q[:,i]=0; Q[i,i]=1.0; c[i]=0.0
    L = np.linalg.cholesky(Q)
    J0 = np.linalg.inv(L).T        # L^-T upper triangular ; x = J0 y
    y = -(J0.T@c)                  # y0 = -L^-1 c
    x = J0@y
    tab = cons_table(mu,fzmin,fzmax)
    Nrm = np.zeros((24,12)); b=np.zeros(24); en=np.zeros(24,bool)
    for j in range(24):
        leg,t=divmod(j,6); ia,ca,ib,cb,bd = tab[t]
        Nrm[j,3*leg+ia]+=ca; Nrm[j,3*leg+ib]+=cb; b[j]=bd; en[j]=bool(contact[leg])
    P=np.eye(12); Ns=np.zeros((12,12)); slot=-np.ones(12,int); u=np.zeros(12); active=np.zeros(24,bool)
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
        nt = J0.T@Nrm[p]           # whitened normal
        z=P@nt; r=Ns@nt
        zeta=nt@z
        dep = zeta <= eps_dep*(nt@nt)
        sp = Nrm[p]@x-b[p]
        act = slot>=0
        t1=np.inf; k=-1
        for kk in range(12):
            if act[kk] and r[kk]>0:
                tt=u[kk]/r[kk]
                if tt<t1: t1=tt;k=kk
        t2 = np.inf if dep else max(0.0,-sp/zeta)
        t=min(t1,t2)
        if t==np.inf: status=2; break
        if not dep:
            y=y+t*z; x = x + t*(J0@z)
        u[act]-=t*r[act]; up+=t
        if t2<=t1:
            q=int(np.argmin(slot>=0))
            P-=np.outer(z,z)/zeta
            Ns[act]-=np.outer(r[act]/zeta,z)
            Ns[q]=z/zeta; slot[q]=p; u[q]=up; active[p]=True; p=-1
        else:
            nu=Ns[k].copy(); delta=nu@nu; gam=Ns@nu
            P+=np.outer(nu,nu)/delta
            Ns-=np.outer(gam/delta,nu)
            Ns[k]=0; active[slot[k]]=False; slot[k]=-1; u[k]=0
            if stats is not None: stats['drops']=stats.get('drops',0)+1
    return status,x,it

if __name__=="__main__":
    p6=default_params(0.6)
    n=int(sys.argv[1]) if len(sys.argv)>1 else 3000
    for masks,seed in (("all4",20260102),("mixed",20260103)):
        for prof in ("default","light","stress"):
            S=states.generate_states(n,seed,profile=prof,masks=masks)
            errs=[]; its=[]; stats={}
            for i in range(len(S)):
                qp=oracle.assemble(p6,S[i:i+1])
                st,xo,lam,ito=oracle.qp_solve(qp['Q'],qp['c'],qp['C'],qp['lb'],qp['ub'])
                st2,x,it=solve(qp['Q'],qp['c'],S['contact'][i],0.6,10.0,120.0,stats=stats)
                assert st2==0,(i,st2)
                errs.append(np.abs(x-xo).max()/max(np.abs(xo).max(),1.0)); its.append(it)
            print(masks,prof,'rel err pct50/99/max',np.percentile(errs,[50,99,100]),'mean it',np.mean(its),'max',max(its),'drops/QP',stats.get('drops',0)/len(S))
