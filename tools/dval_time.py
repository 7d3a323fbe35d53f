"""D-value kernel time on a circuit (tuning aid).  usage: dval_time.py <circuit | path.hgr> ; knobs via the environment."""
import os
import sys
import tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eig_kl_algorithm_b200 import api, datasets
name = sys.argv[1] if len(sys.argv) > 1 else "ibm10"
wd = tempfile.mkdtemp()
if os.path.exists(name):
    path = name
elif name.startswith("synth"):
    path = datasets.write_synthetic(os.path.join(wd, name + ".hgr"), float(name[5:]))
else:
    path = datasets.materialize(wd, circuits=(name,))[name]
with api.Handle() as h:
    h.load_hgr(path); h.assemble_kl_graph()
    n = h.n_nodes
    side = (np.random.default_rng(3).random(n) < 0.5).astype(np.uint8)
    h.set_partition(side)
    warm = h.time_kernel("dvalues", iters=50, flush_l2=False)
    cold = h.time_kernel("dvalues", iters=20, flush_l2=True)
    st = h.stats()
    b = st["bytes_dvalues"]
    print(f"{os.path.basename(path)} carve={os.environ.get('EIGKL_DVAL_CARVE', '-')} dvalues: warm {1e3 * warm:.2f} us ({b / warm / 1e6:.0f} GB/s), "
          f"L2 flushed {1e3 * cold:.2f} us ({b / cold / 1e6:.0f} GB/s) for {b / 1e6:.1f} MB")
