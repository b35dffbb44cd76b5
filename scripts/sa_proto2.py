"""Python SA builder: sign/phase-transport tentative prolongator, strength threshold, dense dofs carried as singletons."""
import numpy as np, scipy.sparse as sp
from sa_proto import rho_dinvA

def aggregate_transport(A, theta, dense):
    n=A.shape[0]; ptr,col,val=A.indptr,A.indices,A.data
    d=np.abs(A.diagonal())
    agg=-np.ones(n,int); tv=np.zeros(n,A.dtype); na=0
    def strong(i,k):
        j=col[k]
        if j==i or dense[j]: return False
        a=abs(val[k]); return a>0 and a*a>=theta*theta*d[i]*d[j]
    ph=lambda v: v/abs(v) if abs(v)>0 else 1.0
    for i in range(n):
        if agg[i]!=-1 or dense[i]: continue
        ks=[k for k in range(ptr[i],ptr[i+1]) if strong(i,k)]
        if not ks or any(agg[col[k]]!=-1 for k in ks): continue
        agg[i]=na; tv[i]=1
        for k in ks:
            j=col[k]; agg[j]=na; tv[j]=ph(-np.conj(val[k]))
        na+=1
    agg1=agg.copy()
    for i in range(n):
        if agg[i]!=-1 or dense[i]: continue
        best=-1;bk=-1
        for k in range(ptr[i],ptr[i+1]):
            if strong(i,k) and agg1[col[k]]!=-1 and abs(val[k])>best: best=abs(val[k]); bk=k
        if bk>=0:
            j=col[bk]; agg[i]=agg1[j]; tv[i]=ph(-val[bk])*tv[j]
    for i in range(n):
        if agg[i]!=-1 or dense[i]: continue
        agg[i]=na; tv[i]=1
        for k in range(ptr[i],ptr[i+1]):
            if strong(i,k) and agg[col[k]]==-1:
                j=col[k]; agg[j]=na; tv[j]=ph(-np.conj(val[k]))
        na+=1
    dense_ids=[]
    for i in np.where(dense)[0]:
        agg[i]=na; tv[i]=1; dense_ids.append(na); na+=1
    return agg,tv,na,dense_ids

def sa_build2(A, theta=0.1, dense_thresh=64, coarse_max=300, omega_scale=4/3, max_levels=12, verbose=True, smooth_dense_cols=True):
    levels=[]; A=sp.csr_matrix(A)
    dense=np.diff(A.indptr)>dense_thresh
    while True:
        n=A.shape[0]; dinv=1/A.diagonal(); lmax=1.1*rho_dinvA(A,dinv)
        if n-dense.sum()<=coarse_max or len(levels)+1>=max_levels:
            levels.append((A,None,lmax)); break
        agg,tv,na,dense_ids=aggregate_transport(A,theta,dense)
        if verbose: print('level',len(levels),'n',n,'agg',na,'ratio %.1f'%(n/na),'dense',int(dense.sum()))
        if na>=0.8*n: levels.append((A,None,lmax)); break
        nrm=np.sqrt(np.bincount(agg,weights=np.abs(tv)**2,minlength=na))
        T=sp.csr_matrix((tv/nrm[agg],(np.arange(n),agg)),shape=(n,na))
        omega=omega_scale/(lmax/1.1)
        AT=(A@T).tocsr()
        if not smooth_dense_cols:
            AT=AT.tolil(); AT[:,dense_ids]=0; AT=AT.tocsr()
        P=(T-omega*sp.diags(dinv)@AT).tocsr()
        Ac=(P.conj().T@A@P).tocsr()
        levels.append((A,P,lmax)); A=Ac
        dense=np.zeros(na,bool); dense[dense_ids]=True
    return levels
