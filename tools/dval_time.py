import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eig_kl_algorithm_b200 import api, datasets
wd = tempfile.mkdtemp()
for name in sys.argv[1:]:
    path = datasets.write_synthetic(os.path.join(wd, name + ".hgr"), float(name[5:])) if name.startswith("synth") else datasets.materialize(wd, circuits=(name,))[name]
    with api.Handle() as h:
        h.load_hgr(path); h.assemble_kl_graph()
        h.set_partition(np.random.default_rng(0).integers(0, 2, h.n_nodes).astype(np.uint8))
        st = h.stats(); b = st["bytes_dvalues"]
        warm = h.time_kernel("dvalues", 50, False); cold = h.time_kernel("dvalues", 20, True)
        print(f"{name:10s} n={st['n_nodes']} nnz={st['nnz_kl']} dvalues warm {warm*1e3:7.2f} us ({b/warm/1e6:6.0f} GB/s) flushed {cold*1e3:7.2f} us ({b/cold/1e6:6.0f} GB/s)")
