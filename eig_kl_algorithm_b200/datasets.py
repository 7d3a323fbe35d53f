"""Input data helpers: unpack the shipped circuits, generate the seeded synthetic circuits.

The reference keeps its data in the CWD (circuit/*.hgr, pre_saved_EIG/<base>_out.txt, results/);
`materialize()` recreates that layout in a working directory from the gzip fixtures under tests/data.
"""
import gzip
import os
import random
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(ROOT, "tests", "data")
REAL_CIRCUITS = ("fract", "ibm01", "industry2", "ibm10")


def _gunzip(src, dst):
    if os.path.exists(dst) and os.path.getmtime(dst) >= os.path.getmtime(src):
        return dst
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    tmp = dst + ".tmp%d" % os.getpid()
    with gzip.open(src, "rb") as fi, open(tmp, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    os.replace(tmp, dst)
    return dst


def materialize(workdir, circuits=REAL_CIRCUITS, golden_eig=True):
    """Unpack circuits (and the reference's golden cEIG outputs) into <workdir>/circuit, /pre_saved_EIG."""
    os.makedirs(os.path.join(workdir, "results"), exist_ok=True)
    os.makedirs(os.path.join(workdir, "pre_saved_EIG"), exist_ok=True)
    out = {}
    for c in circuits:
        out[c] = _gunzip(os.path.join(DATA, "circuit", c + ".hgr.gz"), os.path.join(workdir, "circuit", c + ".hgr"))
        if golden_eig:
            _gunzip(os.path.join(DATA, "pre_saved_EIG", c + ".hgr_out.txt.gz"),
                    os.path.join(workdir, "pre_saved_EIG", c + ".hgr_out.txt"))
    return out


def golden_eig_path(workdir, circuit):
    return os.path.join(workdir, "pre_saved_EIG", circuit + ".hgr_out.txt")


# --------------------------------------------------------------------------------------------------
# Synthetic circuits: same construction as the reference's generator (circuit_generator.py:8-57:
# int(201920*s) nodes, int(210613*s) nets, net size drawn with weights {2:84,3:2,4:6,5:2,6:4,8:2}
# via random.uniform(0,100), pins = sorted random.sample(range(nodes), size) + 1), driven through
# Python's own `random` with an explicit seed (the reference never seeds), so with the same seed the
# output is byte-identical to the reference generator's (checked in tests/test_datasets.py against
# a fixture produced by importing the reference module).
# --------------------------------------------------------------------------------------------------
_SIZES = ((2, 84), (3, 2), (4, 6), (5, 2), (6, 4), (8, 2))


def synthetic_nets(scale, seed=12345):
    rng = random.Random(seed)
    n_nodes = int(201920 * scale)
    n_nets = int(210613 * scale)
    total = sum(p for _, p in _SIZES)
    pop = range(n_nodes)
    nets = []
    uniform, sample = rng.uniform, rng.sample
    for _ in range(n_nets):
        r = uniform(0, total)
        cur = 0
        size = 2
        for s, p in _SIZES:
            cur += p
            if r <= cur:
                size = s
                break
        if size > n_nodes:
            size = n_nodes
        net = sorted(i + 1 for i in sample(pop, size))
        if net:
            nets.append(net)
    return nets, n_nodes


def write_synthetic(path, scale, seed=12345):
    """Writes the seeded synthetic circuit as .hgr (cached: skipped if the file exists)."""
    if os.path.exists(path):
        return path
    nets, n_nodes = synthetic_nets(scale, seed)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    tmp = path + ".tmp%d" % os.getpid()
    with open(tmp, "w") as f:
        f.write(f"{len(nets)} {n_nodes}\n")
        f.write("".join(" ".join(map(str, net)) + "\n" for net in nets))
    os.replace(tmp, path)
    return path


def read_hgr_arrays(path):
    """Host-side .hgr reader for callers of eigkl_set_pins: returns (n_nodes, net_off int64, pins int32).

    Same reading rules as the library's own parser (csrc/hgr_io.cpp): header "<nets> <nodes>", then
    exactly <nets> lines of 1-based ids; missing lines are empty nets.
    """
    import numpy as np
    with open(path, "rb") as f:
        header = f.readline().split()
        n_nets, n_nodes = int(header[0]), int(header[1])
        lines = f.read().split(b"\n")
    lines = lines[:n_nets] + [b""] * max(0, n_nets - len(lines))
    counts = np.fromiter((len(l.split()) for l in lines), dtype=np.int64, count=n_nets)
    net_off = np.zeros(n_nets + 1, np.int64)
    np.cumsum(counts, out=net_off[1:])
    pins = np.array(b" ".join(lines).split(), dtype=np.int64) - 1
    assert len(pins) == net_off[-1]
    return n_nodes, net_off, pins.astype(np.int32)
