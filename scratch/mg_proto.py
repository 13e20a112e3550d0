"""numpy prototype of the mask-aware geometric multigrid used by the CUDA Poisson solver
(design exploration only; not part of the product or the tests)."""
import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from oracle import Oracle, OrcGrid
orc=Oracle()

class Level: pass

def build_levels(M,N,mask,cyl,dx,dz,min_size=4):
    levels=[]
    hx,hz=(dx,dz) if cyl else (1.0,1.0)
    fixed=(mask<2)
    sx=sz=1   # stride of this level's nodes in fine index units
    while True:
        L=Level(); L.M,L.N=M,N; L.fixed=fixed; L.hx,L.hz=hx,hz; L.cyl=cyl; L.sx=sx
        # coefficients: aW (i-1), aE (i+1), aS (j-1), aN(j+1), aC
        i=np.arange(M)[:,None]*np.ones((1,N))
        if cyl:
            ri=i*sx  # radial index in units of fine dx... r = i*hx
            with np.errstate(divide='ignore',invalid='ignore'):
                aW=(i-0.5)/(hx*hx*i); aE=(i+0.5)/(hx*hx*i)
            aS=np.full((M,N),1/(hz*hz)); aN=aS.copy()
            aW[0,:]=0; aE[0,:]=4.0/(hx*hx)
        else:
            aW=np.full((M,N),1/(hx*hx)); aE=aW.copy(); aS=np.full((M,N),1/(hz*hz)); aN=aS.copy()
        aC=-(aW+aE+aS+aN)
        L.a=(aW,aE,aS,aN,aC)
        levels.append(L)
        # choose coarsening: coarsen direction(s) with strongest coupling
        cx = (M-1)//2+1>=min_size and (1/hx**2 >= 0.3/hz**2)
        cz = (N-1)//2+1>=min_size and (1/hz**2 >= 0.3/hx**2)
        if not (cx or cz): break
        fx=2 if cx else 1; fz=2 if cz else 1
        L.fx,L.fz=fx,fz
        Mc=(M-1)//fx+1; Nc=(N-1)//fz+1
        fixed=fixed[::fx,::fz][:Mc,:Nc].copy()
        # if fine grid has odd intervals, the last fine node is dropped: coarse last node keeps its own state
        M,N=Mc,Nc; hx*=fx; hz*=fz; sx*=fx
    return levels

def apply(L,u):
    aW,aE,aS,aN,aC=L.a
    up=np.pad(u,1)
    y=aW*up[:-2,1:-1]+aE*up[2:,1:-1]+aS*up[1:-1,:-2]+aN*up[1:-1,2:]+aC*u
    y[L.fixed]=u[L.fixed]
    return y

def smooth(L,u,b,nsweeps,omega=1.0):
    aW,aE,aS,aN,aC=L.a
    ii,jj=np.indices(u.shape)
    for s in range(nsweeps):
        for color in (0,1):
            up=np.pad(u,1)
            nb=aW*up[:-2,1:-1]+aE*up[2:,1:-1]+aS*up[1:-1,:-2]+aN*up[1:-1,2:]
            new=(b-nb)/aC
            sel=((ii+jj)%2==color)&(~L.fixed)
            u[sel]=(1-omega)*u[sel]+omega*new[sel]
    return u

def restrict(L,Lc,r):
    # full weighting in coarsened directions; residual at fixed nodes is zero
    fx,fz=L.fx,L.fz
    rp=np.pad(r,((1,2),(1,2)))
    M,N=Lc.M,Lc.N
    I=np.arange(M)*fx+1; J=np.arange(N)*fz+1
    def at(di,dj): return rp[np.ix_(I+di,J+dj)]
    if fx==2 and fz==2:
        rc=(4*at(0,0)+2*(at(1,0)+at(-1,0)+at(0,1)+at(0,-1))+at(1,1)+at(1,-1)+at(-1,1)+at(-1,-1))/16
    elif fx==2:
        rc=(2*at(0,0)+at(1,0)+at(-1,0))/4
    else:
        rc=(2*at(0,0)+at(0,1)+at(0,-1))/4
    rc[Lc.fixed]=0
    return rc

def prolong(L,Lc,ec):
    fx,fz=L.fx,L.fz
    e=np.zeros((L.M,L.N))
    Mc,Nc=Lc.M,Lc.N
    # bilinear
    ecp=np.pad(ec,((0,1),(0,1)))
    i=np.arange(L.M); j=np.arange(L.N)
    I=i//fx; J=j//fz
    wi=(i%fx)/fx; wj=(j%fz)/fz
    I1=np.minimum(I+1,Mc); J1=np.minimum(J+1,Nc)
    I=np.minimum(I,Mc); J=np.minimum(J,Nc)   # dropped last node -> pad zero
    e=( (1-wi)[:,None]*(1-wj)[None,:]*ecp[np.ix_(I,J)] + wi[:,None]*(1-wj)[None,:]*ecp[np.ix_(I1,J)]
       +(1-wi)[:,None]*wj[None,:]*ecp[np.ix_(I,J1)] + wi[:,None]*wj[None,:]*ecp[np.ix_(I1,J1)])
    e[L.fixed]=0
    return e

def vcycle(levels,l,u,b,nu1=2,nu2=2):
    L=levels[l]
    if l==len(levels)-1:
        return smooth(L,u,b,50)
    u=smooth(L,u,b,nu1)
    r=b-apply(L,u); r[L.fixed]=0
    Lc=levels[l+1]
    rc=restrict(L,Lc,r)
    ec=vcycle(levels,l+1,np.zeros_like(rc),rc,nu1,nu2)
    u+=prolong(L,Lc,ec)
    return smooth(L,u,b,nu2)

def test(name,g,geo,cyl,seed=0,ncyc=25,probe_radius=1e-4):
    mask,volt=orc.geometry(g,geo,probe_radius,-10.0)
    rng=np.random.default_rng(seed)
    rho=rng.uniform(0,1,(g.M,g.N))*1e-15
    b=orc.rhs(g,mask,volt,rho)
    levels=build_levels(g.M,g.N,mask,cyl,g.dx,g.dz)
    print(name,'levels',[(L.M,L.N) for L in levels])
    u=np.zeros_like(b); u[mask<2]=b[mask<2]
    L0=levels[0]
    # sanity: operator matches oracle
    t=rng.normal(size=b.shape); d=np.abs(apply(L0,t)-orc.apply_operator(g,mask,t)).max(); print('  op check',d/np.abs(t).max())
    r0=None
    for c in range(ncyc):
        u=vcycle(levels,0,u,b)
        r=b-apply(L0,u); r[L0.fixed]=0; rn=np.abs(r).max()
        if r0 is None: r0=rn
        if c%3==0 or c==ncyc-1: print('  cyc',c,'res',rn, 'factor',(rn/r0)**(1/max(c,1)))
        if rn<1e-13*np.abs(b).max(): print('  converged at',c); break
    return u,b,mask

if __name__=='__main__':
    which=sys.argv[1] if len(sys.argv)>1 else 'all'
    if which in('all','box'):
        g=OrcGrid.make(512,512,5.12e-2,5.12e-2,selfconsistent=1,dV=1e-13); test('box512',g,0,False)
    if which in('all','22pt'):
        g=OrcGrid.make(200,200,2e-2,2e-2,rf=1,extern_field=500.0,dV=1e-13); u,b,mask=test('22pt',g,2,False,ncyc=60)
    if which in('all','tube'):
        g=OrcGrid.make(50,50,1.6e-2,1.6e-2,selfconsistent=1,dV=1e-13); test('tube',g,9,False,probe_radius=7.5e-3)
    if which in('all','cyl'):
        g=OrcGrid.make(200,100,1.2e-2,7.5e-2,coord=1,selfconsistent=1,extern_field=500.0,macroparticle_factor=2000); test('cyl empty',g,0,True,ncyc=40)
        g=OrcGrid.make(200,100,1.2e-2,7.5e-2,coord=1,selfconsistent=1,extern_field=500.0,macroparticle_factor=2000); test('cyl penning_simple',g,8,True,ncyc=40)
