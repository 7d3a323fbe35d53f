"""Host emulation of the resident filter's row blocks (first-net order, cost-balanced blocks): how many rows / entries of a block
reference other blocks.  usage: python tools/block_locality_stats.py ibm10 industry2 ibm01  """
import sys, os, gzip, numpy as np, scipy.sparse as sp
sys.path.insert(0, '/root/repo')
from eig_kl_algorithm_b200 import datasets
import tempfile
wd = tempfile.mkdtemp()
for name in sys.argv[1:]:
    path = datasets.materialize(wd, circuits=(name,))[name]
    n_nodes, off, pins = datasets.read_hgr_arrays(path); n_nets = len(off) - 1; pins = pins.astype(np.int64)
    off = np.asarray(off); pins = np.asarray(pins)
    # first-net order
    net_of_pin = np.repeat(np.arange(n_nets), np.diff(off))
    first = np.full(n_nodes, n_nets, dtype=np.int64)
    np.minimum.at(first, pins, net_of_pin)
    perm = np.argsort(first, kind='stable')       # new id -> old id
    inv = np.empty(n_nodes, dtype=np.int64); inv[perm] = np.arange(n_nodes)
    p2 = inv[pins]
    # clique pattern
    rows = []; cols = []
    for e in range(n_nets):
        m = p2[off[e]:off[e+1]]
        k = len(m)
        if k < 2: continue
        a = np.repeat(m, k); b = np.tile(m, k)
        sel = a != b
        rows.append(a[sel]); cols.append(b[sel])
    r = np.concatenate(rows); c = np.concatenate(cols)
    A = sp.coo_matrix((np.ones(len(r)), (r, c)), shape=(n_nodes, n_nodes)).tocsr()
    A.sum_duplicates()
    A = (A + sp.identity(n_nodes, format='csr')).tocsr()   # diagonal included
    A.sort_indices()
    rowptr = A.indptr.astype(np.int64); col = A.indices
    nnz = rowptr[-1]
    G = 148
    cost = nnz + n_nodes
    chunk = max(1024, -(-cost // G)); chunk = -(-chunk // 32) * 32
    nb = max(1, -(-cost // chunk))
    costr = rowptr[:-1] + np.arange(n_nodes)     # cost(r)
    blk = [int(np.searchsorted(costr, b * chunk, side='left')) for b in range(nb)] + [n_nodes]
    out = []
    for b in range(nb):
        r0, r1 = blk[b], blk[b+1]
        if r1 <= r0: continue
        e0, e1 = rowptr[r0], rowptr[r1]
        cc = col[e0:e1]
        remote = (cc < r0) | (cc >= r1)
        rowlen = np.diff(rowptr[r0:r1+1])
        rid = np.repeat(np.arange(r1 - r0), rowlen)
        brow = np.zeros(r1 - r0, bool); brow[np.unique(rid[remote])] = True
        nB = int(rowlen[brow].sum()); nA = int(rowlen[~brow].sum())
        out.append((r1 - r0, e1 - e0, int(brow.sum()), nB, nA, int(remote.sum()), len(np.unique(cc[remote]))))
    out = np.array(out)
    print(name, "blocks", len(out), "nnz", nnz, "chunk", chunk)
    print("  rows/block max", out[:,0].max(), "span max", out[:,1].max(), "halo max", out[:,6].max())
    print("  boundary rows frac: mean %.2f min %.2f max %.2f" % ((out[:,2]/out[:,0]).mean(), (out[:,2]/out[:,0]).min(), (out[:,2]/out[:,0]).max()))
    print("  nB: mean %d max %d ; nA: mean %d max %d ; remote entries frac %.2f" % (out[:,3].mean(), out[:,3].max(), out[:,4].mean(), out[:,4].max(), out[:,5].sum()/out[:,1].sum()))
    for KB, KA in ((8,16),(12,12),(16,8),(20,4)):
        ok = ((out[:,3] <= 512*KB-1) & (out[:,4] <= 512*KA-1)).all()
        print("  split (%d,%d): fits all blocks: %s ; blocks not fitting %d" % (KB, KA, ok, int((~((out[:,3] <= 512*KB-1) & (out[:,4] <= 512*KA-1))).sum())))
