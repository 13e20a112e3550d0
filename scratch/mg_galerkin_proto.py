"""scipy prototype: Galerkin (RAP) multigrid with bilinear P, 4-colour GS, semi-coarsening.
Design exploration for csrc/poisson; not part of the product or tests."""
import sys, numpy as np, scipy.sparse as sp
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from oracle import Oracle, OrcGrid
orc=Oracle()

def fine_matrix(g,mask,cyl):
    M,N=g.M,g.N; n=M*N
    rows=[];cols=[];vals=[]
    fixed=(mask<2).ravel()
    scale=np.ones(n)
    for i in range(M):
        for j in range(N):
            k=i*N+j
            if fixed[k]:
                rows.append(k);cols.append(k);vals.append(1.0);continue
            if cyl:
                k2=1/(g.dz*g.dz)
                if i==0:
                    k3=4/(g.dx*g.dx); ent=[(k-1,k2),(k,-2*k2-k3),(k+1,k2),(k+N,k3)]; s=0.125
                else:
                    k1=(i-0.5)/(g.dx*g.dx*i);k3=(i+0.5)/(g.dx*g.dx*i); ent=[(k-N,k1),(k-1,k2),(k,-2*k2-k1-k3),(k+1,k2),(k+N,k3)]; s=float(i)
                scale[k]=s
            else:
                ent=[(k-N,1.0),(k-1,1.0),(k,-4.0),(k+1,1.0),(k+N,1.0)]
            for c,v in ent:
                if 0<=c<n: rows.append(k);cols.append(c);vals.append(v)
    A=sp.csr_matrix((vals,(rows,cols)),shape=(n,n))
    return A,fixed,scale

def interp_1d(n,f):
    # coarse nodes at fine indices 0,f,2f,...; returns (n x nc) linear interpolation
    if f==1: return sp.identity(n,format='csr'),n
    nc=(n-1)//2+1
    rows=[];cols=[];vals=[]
    for i in range(n):
        I=i//2
        if i%2==0: rows.append(i);cols.append(I);vals.append(1.0)
        else:
            rows.append(i);cols.append(I);vals.append(0.5)
            if I+1<nc: rows.append(i);cols.append(I+1);vals.append(0.5)
    return sp.csr_matrix((vals,(rows,cols)),shape=(n,nc)),nc

def build(g,mask,cyl,min_size=5):
    A,fixed,scale=fine_matrix(g,mask,cyl)
    # symmetrised system S A u = S b, with Dirichlet rows kept as identity and their columns eliminated
    free=(~fixed).astype(float)
    levels=[]
    M,N=g.M,g.N
    hx,hz=(g.dx,g.dz) if cyl else (1.0,1.0)
    S=sp.diags(scale)
    A0=S@A
    Dfree=sp.diags(free)
    # error-equation operator: free rows/cols only + identity on fixed
    Ae=Dfree@A0@Dfree+sp.diags(1.0-free)
    cur=Ae.tocsr(); curfree=free
    while True:
        L=dict(A=cur,M=M,N=N,free=curfree)
        levels.append(L)
        cx=(M-1)//2+1>=min_size and (1/hx**2>=0.3/hz**2)
        cz=(N-1)//2+1>=min_size and (1/hz**2>=0.3/hx**2)
        if not(cx or cz) or len(levels)>12: break
        Px,Mc=interp_1d(M,2 if cx else 1); Pz,Nc=interp_1d(N,2 if cz else 1)
        P=sp.kron(Px,Pz,format='csr')
        P=sp.diags(curfree)@P          # no correction at fixed fine nodes
        Ac=(P.T@cur@P).tocsr()
        d=Ac.diagonal()
        cfree=(np.abs(d)>1e-300).astype(float)
        Ac=Ac+sp.diags(1.0-cfree)
        L['P']=P
        cur=Ac.tocsr(); curfree=cfree; M,N=Mc,Nc; hx*=(2 if cx else 1); hz*=(2 if cz else 1)
    return levels,A,fixed,scale

def smooth(L,u,b,n):
    A=L['A']; M,N=L['M'],L['N']
    d=A.diagonal()
    ii,jj=np.indices((M,N)); col=((ii%2)*2+(jj%2)).ravel()
    for s in range(n):
        for c in range(4):
            r=b-A@u
            sel=(col==c)
            u[sel]+=r[sel]/d[sel]
    return u

def vcycle(levels,l,u,b,nu1=2,nu2=2):
    L=levels[l]
    if l==len(levels)-1:
        return smooth(L,u,b,30)
    u=smooth(L,u,b,nu1)
    r=b-L['A']@u
    rc=L['P'].T@r
    ec=vcycle(levels,l+1,np.zeros_like(rc),rc,nu1,nu2)
    u=u+L['P']@ec
    return smooth(L,u,b,nu2)

def run(name,g,geo,cyl,ncyc=20,probe_radius=1e-4,seed=0,nu=(2,2)):
    mask,volt=orc.geometry(g,geo,probe_radius,-10.0)
    rng=np.random.default_rng(seed)
    rho=rng.uniform(0,1,(g.M,g.N))*1e-15
    b=orc.rhs(g,mask,volt,rho).ravel()
    levels,A,fixed,scale=build(g,mask,cyl)
    print(name,'levels',[(L['M'],L['N']) for L in levels],'nnz/row coarse',[round(L['A'].nnz/L['A'].shape[0],1) for L in levels])
    # u = u_D on fixed; solve for error e: A(uD+e)=b -> Ae e = S(b - A uD) on free
    u=np.zeros_like(b); u[fixed]=b[fixed]
    r0=None
    Ae=levels[0]['A']
    for c in range(ncyc):
        res=scale*(b-A@u); res[fixed]=0
        rn=np.abs(b-A@u)[~fixed].max()
        if r0 is None: r0=rn
        if c%3==0: print('  cyc',c,'res',rn,'factor',(rn/r0)**(1/max(c,1)))
        if rn<1e-13*np.abs(b).max(): print('  converged at',c); break
        e=vcycle(levels,0,np.zeros_like(b),res,*nu)
        u=u+e
    if g.M*g.N<=50000:
        ud=orc.solve_direct(g,mask,b.reshape(g.M,g.N)).ravel(); print('  vs direct',np.abs(u-ud).max()/np.abs(ud).max())

if __name__=='__main__':
    g=OrcGrid.make(512,512,5.12e-2,5.12e-2,selfconsistent=1,dV=1e-13); run('box512',g,0,False,ncyc=13)
    g=OrcGrid.make(200,200,2e-2,2e-2,rf=1,extern_field=500.0,dV=1e-13); run('22pt',g,2,False)
    g=OrcGrid.make(200,200,2e-2,2e-2,rf=1,extern_field=500.0,dV=1e-13); run('8pt',g,3,False)
    g=OrcGrid.make(50,50,1.6e-2,1.6e-2,selfconsistent=1,dV=1e-13); run('tube',g,9,False,probe_radius=7.5e-3)
    g=OrcGrid.make(200,100,1.2e-2,7.5e-2,coord=1,selfconsistent=1,extern_field=500.0,macroparticle_factor=2000); run('cyl empty',g,0,True)
    g=OrcGrid.make(200,100,1.2e-2,7.5e-2,coord=1,selfconsistent=1,extern_field=500.0,macroparticle_factor=2000); run('cyl penning_simple',g,8,True)
    g=OrcGrid.make(101,801,5e-2,45e-2,coord=1,selfconsistent=1,extern_field=500.0,macroparticle_factor=2000); run('cyl MAC',g,6,True)
