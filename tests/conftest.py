import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libomm-lib.so")
PORT_LIB = os.path.join(ROOT, "oracle", "liboracle_port.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: takes more than a few seconds on CPU")


def _load(path):
    from omm_b200.capi import OmmLib
    return OmmLib(path)


@pytest.fixture(scope="session")
def port_lib():
    """oracle/liboracle_port.so -- the plain-C restatement (built on demand: gcc is in the image)."""
    if not os.path.exists(PORT_LIB):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "port"])
    return _load(PORT_LIB)


@pytest.fixture(scope="session")
def ref_lib():
    """oracle/_ref/libomm-lib.so -- the unmodified SDK build (exists where /root/reference was available at build time)."""
    if not os.path.exists(REF_LIB):
        pytest.skip("oracle/_ref/libomm-lib.so not built (needs /root/reference)")
    os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))
    return _load(REF_LIB)


@pytest.fixture(scope="session")
def checker_lib(port_lib):
    """The strongest checker available: the SDK build when present, else the port."""
    if os.path.exists(REF_LIB):
        os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))
        return _load(REF_LIB)
    return port_lib


@pytest.fixture(scope="session")
def product_lib():
    """libomm-b200.so; fails loudly when missing or when no GPU is visible (there is no CPU fallback)."""
    from omm_b200.capi import load_product_library
    lib = load_product_library()
    assert lib.dll.ommB200GetDeviceCount() > 0, "no CUDA device visible: the product cannot run"
    return lib
