import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


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
