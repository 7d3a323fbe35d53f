import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eig_kl_algorithm_b200 import api, datasets
wd = tempfile.mkdtemp()
names = ["ibm01", "industry2", "ibm10"]
paths = datasets.materialize(wd, circuits=tuple(names))
for name in names:
    for ncv in (30, 40, 60, 80, 100, 140):
        for keep in (0,):
            with api.Handle(ncv=ncv, keep=keep) as h:
                h.load_hgr(paths[name]); h.assemble_laplacian()
                h.fiedler(False); lam, _ = h.fiedler(False); st = h.stats()
                print(f"{name:10s} ncv={ncv:3d} matvecs={st['matvecs']:5d} restarts={st['restarts']:3d} fiedler {st['ms_fiedler']:7.2f} ms lambda2={lam:.12g}", flush=True)
