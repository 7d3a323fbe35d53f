"""Halo size of the resident filter's row blocks under different node orders (file order, first-net, reverse Cuthill-McKee,
sorted by the golden Fiedler vector); host emulation.  usage: cd tests && python ../tools/order_locality_stats.py ibm01 ibm10"""
import sys, numpy as np, scipy.sparse as sp, tempfile
from scipy.sparse.csgraph import reverse_cuthill_mckee
sys.path.insert(0, '/root/repo')
from eig_kl_algorithm_b200 import datasets
sys.path.insert(0, '/root/repo/tests')
wd = tempfile.mkdtemp()
def pattern(n_nodes, off, pins):
    rows=[];cols=[]
    for e in range(len(off)-1):
        m=pins[off[e]:off[e+1]]; k=len(m)
        if k<2: continue
        a=np.repeat(m,k); b=np.tile(m,k); sel=a!=b
        rows.append(a[sel]); cols.append(b[sel])
    r=np.concatenate(rows); c=np.concatenate(cols)
    A=sp.coo_matrix((np.ones(len(r)),(r,c)),shape=(n_nodes,n_nodes)).tocsr(); A.sum_duplicates()
    A.data[:]=1
    return A
def stats(A, order, label):
    n=A.shape[0]
    inv=np.empty(n,dtype=np.int64); inv[order]=np.arange(n)
    B=A[order][:,order].tocsr(); B=(B+sp.identity(n,format='csr')).tocsr(); B.sort_indices()
    rowptr=B.indptr.astype(np.int64); col=B.indices
    nnz=rowptr[-1]; G=148; cost=nnz+n
    chunk=max(1024,-(-cost//G)); chunk=-(-chunk//32)*32; nb=max(1,-(-cost//chunk))
    costr=rowptr[:-1]+np.arange(n)
    blk=[int(np.searchsorted(costr,b*chunk,side='left')) for b in range(nb)]+[n]
    H=[];R=[];F=[]
    for b in range(nb):
        r0,r1=blk[b],blk[b+1]
        if r1<=r0: continue
        cc=col[rowptr[r0]:rowptr[r1]]
        rem=(cc<r0)|(cc>=r1)
        H.append(len(np.unique(cc[rem]))); R.append(r1-r0); F.append(rem.mean())
    print("  %-10s halo: mean %5d max %5d | rows mean %d | remote entry frac %.2f" % (label, np.mean(H), np.max(H), np.mean(R), np.mean(F)))
for name in sys.argv[1:]:
    path=datasets.materialize(wd,circuits=(name,))[name]
    n_nodes,off,pins=datasets.read_hgr_arrays(path); pins=pins.astype(np.int64); n_nets=len(off)-1
    A=pattern(n_nodes,off,pins)
    print(name, "n", n_nodes, "nnz", A.nnz+n_nodes)
    net_of_pin=np.repeat(np.arange(n_nets),np.diff(off)); first=np.full(n_nodes,n_nets,dtype=np.int64)
    np.minimum.at(first,pins,net_of_pin)
    stats(A,np.arange(n_nodes),"natural")
    stats(A,np.argsort(first,kind='stable'),"first-net")
    stats(A,np.asarray(reverse_cuthill_mckee(A.tocsr(),symmetric_mode=True)),"RCM")
    try:
        import oracle_lib
        g=oracle_lib.read_eig(datasets.golden_eig_path(wd,name),n_nodes)
        stats(A,np.argsort(g["vec"],kind='stable'),"spectral")
    except Exception as ex:
        print("  spectral: n/a", ex)
