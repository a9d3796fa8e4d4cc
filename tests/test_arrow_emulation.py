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
    lanes=np.arange(NP)
    # ---- elimination: ONE step body for both passes (pass 0: every finger at once, pass 1: the trunk in the same window)
    for ps in (0,1):
        mine=fin if ps==0 else trunk
        fbx=fb if ps==0 else np.zeros(NP,int); fox=fo if ps==0 else np.zeros(NP,int)
        fwx=fw if ps==0 else np.where(trunk,t,0)
        steps=maxw if ps==0 else t; lc0=8*ps
        for s in range(steps):
            pk=fbx+s; act=mine&(s<fwx)
            hk=H[:,0].copy()
            src=np.where(act,pk,lanes)
            dkk=hk[src]
            inv=np.where(act,1/np.sqrt(np.maximum(dkk,1e-20)),1.0)
            lik=hk*inv
            yk=y[src]*inv
            piv=act&(lanes==pk); below=act&(lanes>pk)
            for l in range(NP):
                if piv[l]:
                    myinv[l]=inv[l]; y[l]=yk[l]
                    if ps==0: H[l,8:16]*=inv[l]; tbuf[fox[l]:fox[l]+8]=H[l,8:16]
            for l in range(NP):
                if below[l]: y[l]-=lik[l]*yk[l]; rowb[s&1,fox[l]+l-pk[l]-1]=lik[l]
                Lc[lc0+s,l]=lik[l] if (act[l] and l>=pk[l]) else 0.0
            if below.any():
                for l in range(NP):
                    live=fwx[l]-s-1
                    r=rowb[s&1,fox[l]:fox[l]+8]
                    ml=-lik[l] if below[l] else 0.0
                    newH=H[l].copy()
                    for j in range(7): newH[j]=ml*(r[j] if j<live else 0.0)+H[l,j+1]
                    newH[7]=0
                    if ps==0:
                        v=tbuf[fox[l]:fox[l]+8] if below[l] else np.zeros(8)
                        for cc in range(8): newH[8+cc]=ml*v[cc]+H[l,8+cc]
                    H[l]=newH
        if ps==0:
            # Schur complement by a 4-way split of the finger rows over all lanes, then the trunk rows move into the window
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
            for l in range(t): H[l,0:8]=H[l,8:16]-tot[l,:8]; y[l]-=tot[l,8]
    # ---- back substitution: one step body, trunk pass first, its solution folded into the fingers' right-hand sides
    for ps in (1,0):
        mine=fin if ps==0 else trunk
        fbx=fb if ps==0 else np.zeros(NP,int)
        fwx=fw if ps==0 else np.where(trunk,t,0)
        steps=maxw if ps==0 else t; lc0=8*ps
        if ps==0:
            xt=y[:8].copy()
            for l in range(NP):
                if fin[l]: y[l]-=sum(M[l,cc]*xt[cc] for cc in range(8))
        for s in range(steps-1,-1,-1):
            pk=fbx+s; act=mine&(s<fwx)
            src=np.where(act,pk,lanes); xk=(y*myinv)[src]
            ynew=y.copy()
            for l in range(NP):
                if act[l] and l==pk[l]: ynew[l]=xk[l]
                if act[l] and l<pk[l]: ynew[l]=y[l]-Lc[lc0+l-fbx[l],pk[l]]*xk[l]
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
