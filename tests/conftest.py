import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_cases(basis=None, small_only=False):
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN, "case_*.npz"))):
        name = os.path.basename(f)[5:-4]
        if basis == "newton" and "cheb" in name:
            continue
        if basis == "chebyshev" and "cheb" not in name:
            continue
        if small_only and name in ("c5_d4n20", "c4_d3n30"):
            continue
        out.append(name)
    return out


def load_case(name):
    g = np.load(os.path.join(GOLDEN, f"case_{name}.npz"))
    return g


def case_kwargs(g):
    from lpfun_b200.utils import cheb2nd_nodes, leja_nodes

    return dict(
        spatial_dimension=int(g["m"]),
        polynomial_degree=int(g["n"]),
        lp_degree=float(g["p"]),
        nodes=leja_nodes if str(g["nodes_kind"]) == "leja" else cheb2nd_nodes,
        basis=str(g["basis"]),
        precomputation=bool(g["precomputation"]),
        colex_order=bool(g["colex"]),
    )


@pytest.fixture(scope="session")
def built_lib():
    """Build (if stale) and load liblpfnt.so -- works without a GPU."""
    from lpfun_b200 import _build, _native

    _build.build()
    return _native.lib()
