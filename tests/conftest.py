"""pytest configuration: the `gpu` marker, import paths, shared fixtures.

`-m "not gpu"` covers the oracle against the reference's golden vectors, the host logic, multi-rank sharding on
gloo, and that the C-ABI library loads and exports every symbol include/*.h declares.  `-m gpu` tests are the
parity tests proper: they call the CUDA product through its C-ABI and check it against the oracle.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    from oracle.pyoracle import Port
    return Port()


@pytest.fixture(scope="session")
def ref():
    from oracle.pyoracle import Ref, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return Ref("")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))
