"""One KL pass of a circuit from the reference's shipped EIG partition (ncu target).  usage: kl_one.py <circuit> [passes]"""
import os
import sys
import tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eig_kl_algorithm_b200 import api, datasets  # noqa: E402
name = sys.argv[1] if len(sys.argv) > 1 else "ibm10"
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
wd = tempfile.mkdtemp()
path = datasets.materialize(wd, circuits=(name,))[name]
with api.Handle() as h:
    h.load_hgr(path); h.assemble_kl_graph()
    for _ in range(passes):
        h.load_eig(datasets.golden_eig_path(wd, name)); h.kl_run(False)
    st = h.stats()
    print(name, "swaps", st["kl_swaps"], "loop %.3f ms  %.2f us/swap" % (st["ms_kl_loop"], 1e3 * st["ms_kl_loop"] / max(1, st["kl_swaps"])))
