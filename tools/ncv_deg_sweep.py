import sys, os, tempfile, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from eig_kl_algorithm_b200 import api, datasets
    wd = tempfile.mkdtemp()
    ncv = int(sys.argv[2])
    for name in sys.argv[3:]:
        path = datasets.materialize(wd, circuits=(name,))[name]
        with api.Handle(ncv=ncv) as h:
            h.load_hgr(path); h.assemble_laplacian()
            h.fiedler(False); lam, _ = h.fiedler(False); st = h.stats()
            print(f"deg={os.environ.get('EIGKL_CHEB_DEGREE','dflt'):>4s} ncv={ncv:3d} {name:10s} matvecs={st['matvecs']:5d} steps={st['lanczos_steps']:4d} restarts={st['restarts']:2d} fiedler {st['ms_fiedler']:7.2f} ms  res={st['resid_est'][1]:.2e}", flush=True)
else:
    for d in sys.argv[1].split(","):
        for ncv in sys.argv[2].split(","):
            subprocess.run([sys.executable, __file__, "child", ncv] + sys.argv[3:], env=dict(os.environ, EIGKL_CHEB_DEGREE=d))
