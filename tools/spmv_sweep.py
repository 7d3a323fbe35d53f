"""SpMV / D-value kernel timings per circuit and SpMV mode (tuning aid)."""
import sys, os, tempfile, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    from eig_kl_algorithm_b200 import api, datasets
    name = sys.argv[2]
    wd = tempfile.mkdtemp()
    path = datasets.write_synthetic(os.path.join(wd, name + ".hgr"), float(name[5:])) if name.startswith("synth") else datasets.materialize(wd, circuits=(name,))[name]
    h = api.Handle(); h.load_hgr(path); h.assemble_laplacian(); st = h.stats()
    warm = h.time_kernel("spmv", 50, False); cold = h.time_kernel("spmv", 20, True)
    b = st["bytes_spmv"]
    print(f"{name:10s} mode={os.environ.get('EIGKL_SPMV_MODE','0')} chunk={os.environ.get('EIGKL_CHUNK','auto')} n={st['n_nodes']} nnz={st['nnz_laplacian']} warm {warm*1e3:7.2f} us ({b/warm/1e6:7.0f} GB/s)  flushed {cold*1e3:7.2f} us ({b/cold/1e6:7.0f} GB/s)")
else:
    modes = os.environ.get("MODES", "0,1,2").split(",")
    chunks = os.environ.get("CHUNKS", "").split(",")
    for name in sys.argv[1:]:
        for mode in modes:
            for ch in chunks:
                env = dict(os.environ, EIGKL_SPMV_MODE=mode)
                if ch:
                    env["EIGKL_CHUNK"] = ch
                subprocess.run([sys.executable, __file__, "child", name], env=env)
