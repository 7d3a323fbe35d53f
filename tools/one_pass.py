"""One warm EIG+KL pass of a circuit (ncu target).  usage: one_pass.py <circuit> [warm passes] [keep]"""
import sys, os, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eig_kl_algorithm_b200 import api, datasets
name = sys.argv[1] if len(sys.argv) > 1 else "ibm10"
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 1
keep = int(sys.argv[3]) if len(sys.argv) > 3 else 0
wd = tempfile.mkdtemp()
if name.startswith("synth"):
    path = datasets.write_synthetic(os.path.join(wd, name + ".hgr"), float(name[5:]))
else:
    path = datasets.materialize(wd, circuits=(name,))[name]
prof = os.environ.get("EIGKL_PROFILE") == "1"
h = api.Handle(keep=keep, flags=api.EIGKL_F_PROFILE if prof else 0)
h.load_hgr(path)
for i in range(warm + 1):
    h.invalidate()
    h.assemble_laplacian(); lam, _ = h.fiedler(False); h.partition_from_fiedler(False); h.assemble_kl_graph(); h.kl_run(False)
st = h.stats()
print(name, "lambda2", lam, "matvecs", st["matvecs"], "restarts", st["restarts"], "swaps", st["kl_swaps"], "launches", st["gpu_launches"],
      "ms: fiedler %.2f kl_loop %.2f asmL %.2f asmA %.2f" % (st["ms_fiedler"], st["ms_kl_loop"], st["ms_assemble_laplacian"], st["ms_assemble_kl"]))
if prof:
    per = lambda k: "%s %.2f us x%d" % (k, 1e3 * st["ms_" + k] / max(1, st["n_" + k]), st["n_" + k])
    print("  per launch (all passes):", ", ".join(per(k) for k in ("spmv", "multidot", "update", "restart", "dvalues")))
