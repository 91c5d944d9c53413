"""The C++ ABI driver (charls_b200/csrc/driver) against the unmodified reference library: it binds nothing but the
reference's own entry points, so it must run the reference as well as the B200 library."""
import numpy as np
import pytest

from tests.support import have_reference_build, s_mixed


@pytest.mark.skipif(not have_reference_build(), reason="oracle/_ref/libcharls_ref.so not built")
def test_driver_runs_the_reference_library(reference):
    from charls_b200 import driver

    n, h, w = 7, 40, 120
    frames = np.stack([s_mixed(h, w, 8, seed=i) for i in range(n)])
    streams = np.zeros((n, 2 * h * w + 1024), np.uint8)
    sizes_seen = []
    for per_frame in (True, False):
        out = np.zeros_like(frames)
        seconds, sizes = driver.run_round_trips(reference.path, frames.ctypes.data, h * w, streams.ctypes.data, streams.shape[1],
                                                out.ctypes.data, n, 3, width=w, height=h, bits_per_sample=8, per_frame=per_frame)
        assert seconds > 0 and np.array_equal(out, frames)
        sizes_seen.append(sizes)
    assert sizes_seen[0] == sizes_seen[1]


def test_driver_reports_a_missing_library():
    from charls_b200 import driver

    with pytest.raises(RuntimeError):
        driver.run_round_trips("/nonexistent/libcharls.so", 0, 0, 0, 0, 0, 0, 1, width=1, height=1, bits_per_sample=8)
