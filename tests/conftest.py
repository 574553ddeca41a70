import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name))
    return load


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/oracle.py): the checker, never the thing under test."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as orc
    orc.lib()
    return orc


def ref_available():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "modl", "decomposition", "dict_fact.py"))


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference compiled into oracle/_ref (skips when it was not built)."""
    if not ref_available():
        pytest.skip("oracle/_ref not built (python oracle/build_ref.py)")
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import modl.decomposition.dict_fact as ref_df
    return ref_df


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
