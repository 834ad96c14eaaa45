"""CPU prototype behind the two-level PCG (csrc/cuba_pcg4.cuh): iteration counts of block-Jacobi PCG vs block-Jacobi + coarse
correction over rigid-motion aggregates, on the reduced pose system the CPU oracle assembles.  Test infrastructure (it uses the
oracle); not collected by pytest.  Needs scipy.

    python tests/prototypes/two_level_pcg_prototype.py [kitti00_shaped | ba_kitti_00 | ...]

Output of the run that motivated the design is quoted in profiles/README.md.
"""
import os, sys, time
import numpy as np, scipy.sparse as sp
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
pkg=ge.load_package(); oracle=ge.load_oracle()
name=sys.argv[1] if len(sys.argv)>1 else 'kitti00_shaped'
if name.startswith('ba_'): g=pkg.graphio.read_graph(os.path.join(ROOT,'oracle','_ref','fixtures','%s.cubagraph'%name))
else: g=pkg.synth.make_config(name)
prob=pkg.graphio.flatten(g)
o=oracle.Oracle(prob,(0,0),(0.0,0.0))
o.compute_errors(); o.build_system()
md=o.max_diagonal()
rp,ci=o.hsc_structure()
P=prob.numP
def get(lam):
    assert o.solve(lam)
    Hsc,bsc,inv=o.schur()
    B=Hsc.reshape(-1,6,6).transpose(0,2,1)   # column-major blocks -> [k][r][c]
    rows=np.repeat(np.arange(P),np.diff(rp))
    # assemble full symmetric
    I=[];J=[];V=[]
    for k in range(len(ci)):
        i,j=rows[k],ci[k]
        rr,cc=np.meshgrid(np.arange(6),np.arange(6),indexing='ij')
        I.append((6*i+rr).ravel()); J.append((6*j+cc).ravel()); V.append(B[k].ravel())
        if i!=j:
            I.append((6*j+cc).ravel()); J.append((6*i+rr).ravel()); V.append(B[k].ravel())
    A=sp.csr_matrix((np.concatenate(V),(np.concatenate(I),np.concatenate(J))),shape=(6*P,6*P))
    return A,bsc.reshape(-1).copy()
def pcg(A,b,Minv,tol=1e-11,maxit=20000):
    x=np.zeros_like(b); r=b.copy(); z=Minv(r); p=z.copy(); rz=r@z; rz0=rz; it=0
    while it<maxit:
        Ap=A@p; al=rz/(p@Ap); x+=al*p; r-=al*Ap; z=Minv(r); rzn=r@z; it+=1
        if rzn<=tol*tol*rz0: break
        p=z+(rzn/rz)*p; rz=rzn
    return x,it
def block_jacobi(A):
    D=[np.linalg.inv(A[6*i:6*i+6,6*i:6*i+6].toarray()) for i in range(P)]
    Dm=sp.block_diag(D,format='csr')
    return lambda r: Dm@r
def two_level(A,m,adj=None):
    bj=block_jacobi(A)
    na=(P+m-1)//m
    # Z: piecewise constant 6-dof per aggregate (optionally adjoint-transformed)
    I=[];J=[];V=[]
    for i in range(P):
        a=i//m
        Zi=np.eye(6) if adj is None else adj[i]
        rr,cc=np.meshgrid(np.arange(6),np.arange(6),indexing='ij')
        I.append((6*i+rr).ravel()); J.append((6*a+cc).ravel()); V.append(Zi.ravel())
    Z=sp.csr_matrix((np.concatenate(V),(np.concatenate(I),np.concatenate(J))),shape=(6*P,6*na))
    Ac=(Z.T@A@Z).toarray(); Aci=np.linalg.inv(Ac)
    return lambda r: bj(r)+Z@(Aci@(Z.T@r)), na
def adjoints():
    # pose update T <- Exp(delta) T with delta in camera frame; world-frame rigid twist xi=(w,v): camera i sees delta_i = Ad(T_i) xi
    q=prob.q[:P]; t=prob.t[:P]
    out=[]
    for i in range(P):
        x,y,z,w=q[i]
        R=np.array([[1-2*(y*y+z*z),2*(x*y-z*w),2*(x*z+y*w)],[2*(x*y+z*w),1-2*(x*x+z*z),2*(y*z-x*w)],[2*(x*z-y*w),2*(y*z+x*w),1-2*(x*x+y*y)]])
        tx=np.array([[0,-t[i][2],t[i][1]],[t[i][2],0,-t[i][0]],[-t[i][1],t[i][0],0]])
        Ad=np.zeros((6,6)); Ad[:3,:3]=R; Ad[3:,3:]=R; Ad[3:,:3]=tx@R
        out.append(Ad)
    return out
adj=adjoints()
for lam in (1e-5*md, 1e-5*md/1e3, 1e-5*md/1e5):
    A,b=get(lam)
    t0=time.time(); x,it=pcg(A,b,block_jacobi(A)); 
    print(name,'lambda %.3g'%lam,'block-Jacobi iters',it,flush=True)
    for m in (8,16,32):
        M,na=two_level(A,m); x2,it2=pcg(A,b,M)
        M3,_=two_level(A,m,adj); x3,it3=pcg(A,b,M3)
        print('   two-level m=%d (nc=%d): piecewise-const iters %d   adjoint(rigid) iters %d   |dx| rel %.1e'%(m,6*na,it2,it3,np.abs(x3-x).max()/np.abs(x).max()),flush=True)
