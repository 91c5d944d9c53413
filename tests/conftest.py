import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _cuda_device_count():
    """Devices the in-tree library sees (0 when it is not built, no driver, or no GPU)."""
    try:
        import ctypes

        from charls_b200 import capi

        if not os.path.exists(capi.DEFAULT_LIBRARY):
            return 0
        n = ctypes.c_int32(0)
        if capi.default_library().charlsx_get_device_count(ctypes.byref(n)) != 0:
            return 0
        return n.value
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """A plain `pytest tests` on a host without CUDA skips the gpu-marked tests instead of failing in them (the error path
    itself -- compute calls fail loudly with error 200 -- has its own unmarked test in tests/test_abi.py).  With `-m gpu`
    asked for explicitly nothing is skipped: on the GPU box a missing device must be a failure, not a green run."""
    if "gpu" in (config.getoption("-m") or ""):
        return
    if _cuda_device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device (gpu-marked tests run with -m gpu on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from tests.support import oracle as make

    return make()


@pytest.fixture(scope="session")
def reference():
    """The unmodified reference build (oracle/_ref); built on demand where /root/reference exists, else skipped."""
    from tests import support

    if not support.have_reference_build() and not os.path.isdir(support.REFERENCE_ROOT):
        pytest.skip("oracle/_ref/libcharls_ref.so not built")
    return support.reference_library()


@pytest.fixture(scope="session")
def product():
    """The in-tree CUDA build of the ABI.  Fails (does not skip) when it is missing: there is no fallback."""
    from charls_b200 import capi

    if not os.path.exists(capi.DEFAULT_LIBRARY):
        import __graft_entry__

        __graft_entry__.build()
    return capi.default_library()


@pytest.fixture(scope="session")
def golden():
    from tests.golden_vectors import load_vectors

    return load_vectors()
