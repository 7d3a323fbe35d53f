"""The clique-model Laplacian (cEIG.cpp:86-133: weight 2/k per pin pair of a k-pin net) of a shipped circuit, scipy CSR."""
import gzip
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as sla

DATA = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), 'tests', 'data', 'circuit')


def load_L(name):
    f = gzip.open(os.path.join(DATA, name + '.hgr.gz'), 'rt').read().split('\n')
    nn,N=map(int,f[0].split()[:2])
    r=[];c=[];v=[]
    for line in f[1:1+nn]:
        p=np.array([int(x)-1 for x in line.split()])
        k=len(p)
        if k<2: continue
        w=2.0/k
        ii,jj=np.meshgrid(p,p)
        m=ii!=jj
        r.append(ii[m]);c.append(jj[m]);v.append(np.full(m.sum(),w))
    r=np.concatenate(r);c=np.concatenate(c);v=np.concatenate(v)
    A=sp.csr_matrix((v,(r,c)),shape=(N,N)); A.sum_duplicates()
    d=np.asarray(A.sum(axis=1)).ravel()
    return (sp.diags(d)-A).tocsr(), d
if __name__=='__main__':
    name=sys.argv[1]
    L,d=load_L(name)
    t=time.time()
    k=40
    vals=sla.eigsh(L,k=k,sigma=-1e-3,which='LM',return_eigenvectors=False,tol=1e-9)
    vals=np.sort(vals)
    print(name,'n',L.shape[0],'nnz',L.nnz,'diag min/max',d.min(),d.max(),'time',round(time.time()-t,1))
    print('lowest eigenvalues:',np.array2string(vals[:k],precision=5,max_line_width=200))
