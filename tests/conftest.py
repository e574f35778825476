import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle

    oracle.build()
    return oracle


def make_basis(orc, name):
    """-> (geometry, shells, oracle Basis, nocc) for a named workload."""
    from quantumlab_jl_b200 import workloads
    from quantumlab_jl_b200.shell import flattenShells

    geo, shells, nocc = workloads.build(name)
    return geo, shells, orc.Basis(*flattenShells(shells)), nocc


def atoms_of(geo):
    import numpy as np

    xyz = np.array([a.position for a in geo.atoms], dtype=float)
    Z = np.array([a.element.atomicNumber for a in geo.atoms], dtype=float)
    return xyz, Z
