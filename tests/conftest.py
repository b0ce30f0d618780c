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


def _gpu_usable() -> bool:
    """fvs2d_gpu_init succeeds only with a usable CUDA device (the library has no CPU fallback)."""
    try:
        from fvs2d_b200 import capi, config
        L = capi.lib()
        import ctypes
        cfg = config.RunInput(lvortex=True).to_config()
        ok = L.fvs2d_gpu_init(ctypes.byref(cfg), 0) == 0
        L.fvs2d_gpu_finalize()
        return ok
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """A plain `pytest` on a box without a GPU skips the gpu-marked tests instead of failing them (the product itself still
    fails loudly there: tests/test_host_logic.py::test_no_gpu_means_loud_failure)."""
    if not any("gpu" in it.keywords for it in items) or _gpu_usable():
        return
    skip = pytest.mark.skip(reason="no usable CUDA device: gpu-marked tests run on the B200 box (pytest -m gpu)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


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
