import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """The CPU restatement (oracle/libpm_oracle.so); built on demand with gcc."""
    from oracle import oraclelib
    if not oraclelib.available():
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libpm_oracle.so"])
    return oraclelib.Oracle()


@pytest.fixture(scope="session")
def refhost():
    """The reference's own routines compiled as host code (oracle/_ref); absent on a fresh clone without /root/reference."""
    from oracle import refhost as rh
    if not rh.available():
        if os.path.exists("/root/reference/photonMappingKernel.cu"):
            subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")])
        if not rh.available():
            pytest.skip("oracle/_ref/libpmref_host.so not built and /root/reference absent")
    return rh.RefHost()


@pytest.fixture(scope="session")
def pm():
    """The product module (ctypes over libpmb200.so)."""
    import pmb200
    pmb200.lib()
    return pmb200
