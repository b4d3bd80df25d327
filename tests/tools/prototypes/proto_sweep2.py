"""Two pivots per barrier interval in the MPC Gauss-Jordan sweep: numpy emulation of the per-element
rules of the CUDA experiment (tile masks included) against the one-pivot sweep and inv(cholesky(H))."""
import numpy as np
def tiles_mask(n, NB, BJ):
    # element (R,C) updated iff C in blocks >= BJ and (a <= BJ or b <= a)
    R=np.arange(16*NB)[:,None]; C=np.arange(16*NB)[None,:]
    a=R//16; b=C//16
    return (b>=BJ) & ((a<=BJ)|(b<=a))
def single(H):
    n=H.shape[0]; NB=2*((n+31)//32); N=16*NB
    m=np.zeros((N,N)); m[:n,:n]=H
    dval=np.zeros(n)
    for j in range(n):
        BJ=j//16; tj=j%16
        c=m[:,j].copy(); d=c[j]; rd=1/d; dval[j]=d
        cv=c.copy(); cv[j]=0
        ci=c*rd; ci[:j+1]=0   # columns <= j finished (within block: tx<=tj; blocks < BJ not looped)
        mask=tiles_mask(n,NB,BJ)
        upd=np.outer(cv,ci)
        m=np.where(mask, m-upd, m)
        m[j,j+1:]=-ci[j+1:]
    return m,dval
def fused(H):
    n=H.shape[0]; NB=2*((n+31)//32); N=16*NB
    m=np.zeros((N,N)); m[:n,:n]=H
    dval=np.zeros(n)
    j=0
    while j<n:
        BJ=j//16
        if j+1>=n:
            c=m[:,j].copy(); d=c[j]; rd=1/d; dval[j]=d
            cv=c.copy(); cv[j]=0; ci=c*rd; ci[:j+1]=0
            mask=tiles_mask(n,NB,BJ)
            m=np.where(mask,m-np.outer(cv,ci),m); m[j,j+1:]=-ci[j+1:]
            break
        c1=m[:,j].copy(); c2=m[:,j+1].copy()
        d1=c1[j]; e=c1[j+1]; g=c2[j+1]; rd1=1/d1; l=e*rd1; d2=g-e*l; rd2=1/d2
        dval[j]=d1; dval[j+1]=d2
        cv1=c1.copy(); cv2=c2-l*c1
        cv1[j]=0; cv2[j]=-l; cv2[j+1]=0
        ci1=c1*rd1; ci2=(c2-l*c1)*rd2
        ci1[:j+1]=0; ci2[:j+2]=0
        mask=tiles_mask(n,NB,BJ)
        m=np.where(mask, m-np.outer(cv1,ci1)-np.outer(cv2,ci2), m)
        # row assignments
        cols=np.arange(N)
        r1=cols>j; m[j,r1]=l*ci2[r1]-ci1[r1]
        r2=cols>j+1; m[j+1,r2]=-ci2[r2]
        j+=2
    return m,dval
rng=np.random.default_rng(0)
for n in (3,6,15,16,17,33,48,63,90,119,120):
    A=rng.normal(size=(n,n)); H=A@A.T+n*np.eye(n)
    m1,d1=single(H); m2,d2=fused(H)
    iu=np.triu_indices(n,1)
    X1=m1[:n,:n][iu]; X2=m2[:n,:n][iu]
    # reference: X = L^-1 ; output X[i][c] = m[c][i]/sqrt(d_i) for c<i
    L=np.linalg.cholesky(H); X=np.linalg.inv(L)
    Xg=np.zeros((n,n))
    for c in range(n):
        for i in range(c+1,n): Xg[i,c]=m2[c,i]/np.sqrt(d2[i])
    err=np.abs(Xg-np.tril(X,-1)).max()
    print(n, 'single vs fused', np.abs(X1-X2).max(), 'dval', np.abs(d1-d2).max(), 'vs true inverse', err, 'diag', np.abs(1/np.sqrt(d2)-np.diag(X)).max())
