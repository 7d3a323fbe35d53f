import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")
CIRCUITS = ("fract", "ibm01", "industry2", "ibm10")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def workdir(tmp_path_factory):
    """A directory laid out like the reference's CWD: circuit/, pre_saved_EIG/ (golden), results/."""
    from eig_kl_algorithm_b200 import datasets
    d = str(tmp_path_factory.mktemp("eigkl_work"))
    datasets.materialize(d)
    return d


@pytest.fixture(scope="session")
def circuits(workdir):
    return {c: os.path.join(workdir, "circuit", c + ".hgr") for c in CIRCUITS}


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def eigkl_lib():
    """Builds (if needed) and loads libeigkl.so.  Loading needs no GPU."""
    from eig_kl_algorithm_b200 import build as _build, api
    _build.build()
    return api.load_library()
