"""KL swap-loop time vs cluster size (tuning aid)."""
import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eig_kl_algorithm_b200 import api, datasets
wd = tempfile.mkdtemp()
names = sys.argv[1:] or ["ibm01", "industry2", "ibm10"]
paths = datasets.materialize(wd, circuits=tuple(names))
for name in names:
    for nc in (1, 2, 4, 8, 16):
        with api.Handle(kl_cluster=nc) as h:
            h.load_hgr(paths[name]); h.assemble_kl_graph()
            best = 1e9
            for rep in range(3):
                h.load_eig(datasets.golden_eig_path(wd, name)); h.kl_run(False); st = h.stats(); best = min(best, st["ms_kl_loop"])
            print(f"{name:10s} cluster={nc:2d} swaps={st['kl_swaps']} loop {best:8.3f} ms  {1e3*best/st['kl_swaps']:6.2f} us/swap  setup {st['ms_kl_setup']:.2f} ms", flush=True)
