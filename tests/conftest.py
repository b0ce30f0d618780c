import json
import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The oracle is compiled on demand (seconds); the product library must already exist."""
    from oracle import oracle
    oracle.build()


def run_input(name: str):
    from fvs2d_b200 import config
    d = json.load(open(os.path.join(GOLDEN, "inputs.json")))[name]
    return config.RunInput(**{k: (tuple(v) if isinstance(v, list) else v) for k, v in d.items()})


@pytest.fixture(scope="session")
def vortex_mesh():
    from fvs2d_b200 import meshio
    return meshio.load_npz(os.path.join(GOLDEN, "vortex_mesh.npz"))


@pytest.fixture(scope="session")
def naca_mesh():
    from fvs2d_b200 import meshio
    return meshio.load_npz(os.path.join(GOLDEN, "naca_mesh.npz"))
