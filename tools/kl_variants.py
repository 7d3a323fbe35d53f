"""KL swap-loop variants side by side (tuning aid): us per swap of the flat / warp-per-row shared-memory loops and the
cluster loop, with the in-kernel phase clocks of the shared-memory forms.

    kl_variants.py [circuit ...]          circuits: ibm01 industry2 ibm10 synth<scale>

Real circuits start from the reference's shipped EIG partition, synthetic ones from this library's own Fiedler split.
Every variant runs in a fresh handle (the knobs are read when a handle is created)."""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eig_kl_algorithm_b200 import api, datasets  # noqa: E402

names = sys.argv[1:] or ["ibm10", "synth1"]
wd = tempfile.mkdtemp()
VARIANTS = [("flat", {"EIGKL_KL_FLAT": "1"}), ("rows", {"EIGKL_KL_FLAT": "0"}), ("cluster", {"EIGKL_KL_LOCAL": "0"})]
if os.environ.get("KLV_ONLY"):
    VARIANTS = [v for v in VARIANTS if v[0] in os.environ["KLV_ONLY"].split(",")]
for name in names:
    if name.startswith("synth"):
        path = datasets.write_synthetic(os.path.join(wd, name + ".hgr"), float(name[5:]))
        with api.Handle() as h:
            h.load_hgr(path); h.assemble_laplacian(); h.fiedler(False)
            _, side = h.partition_from_fiedler()
    else:
        path = datasets.materialize(wd, circuits=(name,))[name]
        side = None
    ref = None
    for label, env in VARIANTS:
        for phases in (False, True):
            for k in ("EIGKL_KL_FLAT", "EIGKL_KL_LOCAL", "EIGKL_KL_PHASES"):
                os.environ.pop(k, None)
            os.environ.update(env)
            if phases:
                if label == "cluster":
                    continue
                os.environ["EIGKL_KL_PHASES"] = "1"
            with api.Handle() as h:
                h.load_hgr(path); h.assemble_kl_graph()
                best = 1e30
                for rep in range(1 if phases else 3):
                    if side is None:
                        h.load_eig(datasets.golden_eig_path(wd, name))
                    else:
                        h.set_partition(side)
                    tr = h.kl_run()
                    st = h.stats()
                    best = min(best, st["ms_kl_loop"])
                sig = (tr["swaps"], tr["node1"].tobytes(), tr["cut"].tobytes())
                if ref is None:
                    ref = sig
                print(f"{name:10s} {label:8s}{' +clocks' if phases else '        '} swaps={st['kl_swaps']:7d} loop {best:9.3f} ms "
                      f"{1e3 * best / max(1, st['kl_swaps']):6.2f} us/swap  setup {st['ms_kl_setup']:.2f} ms  same_trace={sig == ref}", flush=True)
