import importlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "needs_reference: executes the unmodified reference from /root/reference")


@pytest.fixture(scope="session", autouse=True)
def _built_oracle():
    lib = os.path.join(ROOT, "oracle", "_build", "liboracle.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    yield


@pytest.fixture(scope="session")
def amp():
    """The product package (3d-semantic-segmentation-amp-net_b200/)."""
    return importlib.import_module("3d-semantic-segmentation-amp-net_b200")


@pytest.fixture(scope="session")
def reference():
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not present (GPU box)")
    return ref_import.load()


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
