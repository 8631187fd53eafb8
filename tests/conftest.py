import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _ensure_built():
    """The shared libraries are build artefacts (git-ignored): compile them when a checkout has none
    (nvcc cross-compiles sm_100a without a GPU; the recipe is __graft_entry__.build's)."""
    import subprocess
    lib = os.path.join(ROOT, "ocaml-hnsw_b200", "libhnsw_b200.so")
    orc = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "ocaml-hnsw_b200"), "libhnsw_b200.so"])
    if not os.path.exists(orc):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])


def pytest_configure(config):
    _ensure_built()
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: larger CPU cases")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
