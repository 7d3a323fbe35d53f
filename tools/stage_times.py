"""Wall-clock per API call, resident vs e2e path (debug aid)."""
import sys, time, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from eig_kl_algorithm_b200 import api, datasets
name = sys.argv[1] if len(sys.argv) > 1 else "ibm10"
wd = tempfile.mkdtemp()
path = datasets.materialize(wd, circuits=(name,))[name]
n, off, pins = datasets.read_hgr_arrays(path)
h = api.Handle()
def T(f, *a, **k):
    t = time.perf_counter(); r = f(*a, **k); return (time.perf_counter() - t) * 1e3, r
for mode in ("e2e", "resident", "resident", "e2e", "resident"):
    row = []
    if mode == "e2e": row.append(("set_pins", T(h.set_pins, n, off, pins)[0]))
    else: row.append(("invalidate", T(h.invalidate)[0]))
    row.append(("asm_L", T(h.assemble_laplacian)[0]))
    row.append(("fiedler", T(h.fiedler, False)[0]))
    row.append(("partition", T(h.partition_from_fiedler, False)[0]))
    row.append(("asm_A", T(h.assemble_kl_graph)[0]))
    row.append(("kl", T(h.kl_run, False)[0]))
    st = h.stats()
    print(mode, " ".join(f"{k}={v:.1f}" for k, v in row), "| dev: asmL=%.1f asmA=%.1f fied=%.1f klsetup=%.1f klloop=%.1f" % (
        st["ms_assemble_laplacian"], st["ms_assemble_kl"], st["ms_fiedler"], st["ms_kl_setup"], st["ms_kl_loop"]), flush=True)
