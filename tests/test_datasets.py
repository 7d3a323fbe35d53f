"""Synthetic-circuit generator restatement and the host-side .hgr reader."""
import os

import numpy as np

from conftest import GOLDEN
from eig_kl_algorithm_b200 import datasets


def test_synthetic_matches_reference_generator(tmp_path):
    # fixture produced by importing the reference's circuit_generator.py (tests/golden/make_synth_golden.py)
    out = str(tmp_path / "s.hgr")
    datasets.write_synthetic(out, 0.002, seed=12345)
    assert open(out, "rb").read() == open(os.path.join(GOLDEN, "synth_0.002_seed12345.hgr"), "rb").read()


def test_synthetic_scale_0_1_statistics(tmp_path):
    # SURVEY.md Appendix F, seed 12345, scale 0.1: 21 061 nets, 20 192 nodes, 52 230 pins, 55 067 clique pairs
    out = str(tmp_path / "s.hgr")
    datasets.write_synthetic(out, 0.1, seed=12345)
    n, off, pins = datasets.read_hgr_arrays(out)
    k = np.diff(off)
    assert (n, len(off) - 1, len(pins), int((k * (k - 1) // 2).sum())) == (20192, 21061, 52230, 55067)
    assert pins.min() >= 0 and pins.max() < n
    for e in range(0, len(k), 997):                      # pins sorted ascending and distinct inside a net
        p = pins[off[e]:off[e + 1]]
        assert np.all(np.diff(p) > 0)


def test_read_hgr_arrays_matches_oracle(oracle, circuits):
    for c in ("fract", "ibm01"):
        n, off, pins = datasets.read_hgr_arrays(circuits[c])
        h = oracle.OracleHgr(circuits[c])
        assert n == h.n_nodes and np.array_equal(off, h.net_off) and np.array_equal(pins, h.pins)
