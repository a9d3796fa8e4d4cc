"""Lane-by-lane numpy emulation of the arrow factorisation in dexr_kernels.cuh (Solver<32, -1>): the same registers per lane
(8 own-finger columns rotating, 8 trunk columns), the same scratch buffers and read/write pattern (poisoned with NaN here: every
read of a slot nobody wrote must be masked by a select, never by a multiplication by zero), the same order of operations.
It checks the ALGEBRA -- fingers side by side, Schur complement by a 4-way split, trunk, back substitution -- against
numpy.linalg.solve on random symmetric positive definite arrow matrices; the CUDA code itself is checked on the GPU
(tests/test_gpu_arrow.py)."""
import numpy as np
import pytest

rng = np.random.RandomState(0)
NP=32
def run(t, widths, dof):
    # structure
    fb=np.zeros(NP,int); fw=np.zeros(NP,int); fo=np.zeros(NP,int); fin=np.zeros(NP,bool)
    c=t
    for f,w in enumerate(widths):
        for i in range(c,c+w): fb[i]=c; fw[i]=w; fo[i]=8*(1+f); fin[i]=True
        c+=w
    assert c==dof
    trunk=np.arange(NP)<t
    maxw=max(widths)
    # random SPD arrow matrix
    A=np.zeros((dof,dof))
    J=rng.randn(60,dof)
    # residual rows touch trunk + one finger
    c=t
    Jm=np.zeros_like(J)
    starts=[t+sum(widths[:f]) for f in range(len(widths))]
    for r in range(60):
        f=r%len(widths); Jm[r,:t]=J[r,:t]; Jm[r,starts[f]:starts[f]+widths[f]]=J[r,starts[f]:starts[f]+widths[f]]
    A=Jm.T@Jm+0.1*np.eye(dof)
    g=rng.randn(dof)
    ref=np.linalg.solve(A,-g)
    # lane registers
    H=np.zeros((NP,16)); y=np.zeros(NP); myinv=np.ones(NP)
    for l in range(dof):
        if fin[l]:
            for j in range(fw[l]): H[l,j]=A[l,fb[l]+j]
            for cc in range(t): H[l,8+cc]=A[l,cc]
        else:
            for d in range(t): H[l,8+d]=A[l,d]
        y[l]=-g[l]
    rowb=np.full((2,64),np.nan); tbuf=np.full(64,np.nan); Lc=np.full((16,NP+1),np.nan); M=np.zeros((NP,12))
    for s in range(maxw):
        pk=fb+s; act=fin&(s<fw)
        hk=H[:,0].copy()
        src=np.where(act,pk,np.arange(NP))
        dkk=hk[src]
        inv=np.where(act,1/np.sqrt(np.maximum(dkk,1e-20)),1.0)
        lik=hk*inv
        yk=y[src]*inv
        piv=act&(np.arange(NP)==pk); below=act&(np.arange(NP)>pk)
        for l in range(NP):
            if piv[l]:
                myinv[l]=inv[l]; y[l]=yk[l]; H[l,8:16]*=inv[l]; tbuf[fo[l]:fo[l]+8]=H[l,8:16]
        for l in range(NP):
            if below[l]: y[l]-=lik[l]*yk[l]; rowb[s&1,fo[l]+l-pk[l]-1]=lik[l]
            Lc[s,l]=lik[l] if (act[l] and l>=pk[l]) else 0.0
        if below.any():
            for l in range(NP):
                live=fw[l]-s-1
                r=rowb[s&1,fo[l]:fo[l]+8]
                v=tbuf[fo[l]:fo[l]+8] if below[l] else np.zeros(8)
                ml=-lik[l] if below[l] else 0.0
                newH=H[l].copy()
                for j in range(7): newH[j]=ml*(r[j] if j<live else 0.0)+H[l,j+1]
                newH[7]=0
                for cc in range(8): newH[8+cc]=ml*v[cc]+H[l,8+cc]
                H[l]=newH
    for l in range(NP):
        M[l,8]=y[l] if fin[l] else 0
        M[l,:8]=H[l,8:16] if fin[l] else 0
    acc=np.zeros((NP,9))
    for l in range(NP):
        c_=l&7; qd=l>>3
        for i in range(t+qd,dof,4):
            acc[l,:8]+=M[i,c_]*M[i,:8]; acc[l,8]+=M[i,c_]*M[i,8]
    tot=np.zeros((NP,9))
    for l in range(NP):
        for q in range(4): tot[l]+=acc[(l&7)+8*q]
    for l in range(t): H[l,8:16]-=tot[l,:8]; y[l]-=tot[l,8]
    for k in range(t):
        hk=H[:,8].copy(); dkk=hk[k]; inv=1/np.sqrt(max(dkk,1e-20)); lik=hk*inv; yk=y[k]*inv
        for l in range(NP):
            below=trunk[l] and l>k
            if l==k: myinv[l]=inv; y[l]=yk
            if below: y[l]-=lik[l]*yk; rowb[k&1,l-k-1]=lik[l]
            Lc[8+k,l]=lik[l]
        live=t-k-1
        for l in range(NP):
            below=trunk[l] and l>k
            ml=-lik[l] if below else 0.0
            r=rowb[k&1,0:8]
            newH=H[l].copy()
            for j in range(7): newH[8+j]=ml*(r[j] if j<live else 0.0)+H[l,9+j]
            newH[15]=0; H[l]=newH
    for k in range(t-1,-1,-1):
        xk=y[k]*myinv[k]
        for l in range(NP):
            if l==k: y[l]=xk
            if trunk[l] and l<k: y[l]-=Lc[8+l,k]*xk
    for l in range(NP):
        if fin[l]: y[l]-=sum(M[l,cc]*y[cc] for cc in range(8))
    for s in range(maxw-1,-1,-1):
        pk=fb+s; act=fin&(s<fw)
        src=np.where(act,pk,np.arange(NP)); xk=(y*myinv)[src]
        ynew=y.copy()
        for l in range(NP):
            if act[l] and l==pk[l]: ynew[l]=xk[l]
            if act[l] and l<pk[l]: ynew[l]=y[l]-Lc[l-fb[l],pk[l]]*xk[l]
        y=ynew
    return np.abs(y[:dof]-ref).max()


@pytest.mark.parametrize("t,widths", [(8, [4, 5, 4, 4, 5]),     # Shadow on a free-flying base
                                      (2, [4, 5, 4, 4, 5]),     # Shadow teleop
                                      (6, [4, 4, 4, 4]),        # Allegro / LEAP on a free-flying base
                                      (0, [5, 5, 5, 5]),        # no trunk at all
                                      (1, [8, 8, 8, 7]),        # widest fingers, all 32 lanes
                                      (8, [1, 1, 8, 2, 3, 6])]) # six ragged fingers
def test_arrow_factorisation_solves_the_system(t, widths):
    assert run(t, widths, t + sum(widths)) < 1e-12
