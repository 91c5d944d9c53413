"""ctypes wrapper of charls_b200/lib/libabi_driver.so: image round trips through a CharLS-compatible C ABI from C++ threads.

One call runs `n` encode (+ decode) round trips with `threads` worker threads inside the driver, so the interpreter lock
is not part of what gets timed.  The driver binds the library by path with dlopen, which makes it usable with the B200
library and with the reference's libcharls alike (bench.py uses it for both arms)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DRIVER_LIBRARY = os.path.join(_HERE, "lib", "libabi_driver.so")


class Job(C.Structure):
    _fields_ = [
        ("library_path", C.c_char_p), ("frames", C.c_void_p), ("frame_bytes", C.c_size_t), ("streams", C.c_void_p),
        ("stream_capacity", C.c_size_t), ("decoded", C.c_void_p), ("sizes", C.POINTER(C.c_size_t)), ("n", C.c_int32),
        ("threads", C.c_int32), ("width", C.c_uint32), ("height", C.c_uint32), ("bits_per_sample", C.c_int32),
        ("component_count", C.c_int32), ("near_lossless", C.c_int32), ("interleave_mode", C.c_int32),
        ("color_transformation", C.c_int32), ("per_frame", C.c_int32), ("seconds", C.c_double),
    ]


_dll = None


def _driver():
    global _dll
    if _dll is None:
        if not os.path.exists(DRIVER_LIBRARY):
            raise FileNotFoundError(f"{DRIVER_LIBRARY} is missing: run python -m charls_b200.build")
        _dll = C.CDLL(DRIVER_LIBRARY)
        _dll.abi_driver_run.restype = C.c_int32
        _dll.abi_driver_run.argtypes = [C.POINTER(Job)]
    return _dll


def run_round_trips(library_path, frames_ptr, frame_bytes, streams_ptr, stream_capacity, decoded_ptr, n, threads, *, width, height,
                    bits_per_sample, component_count=1, near_lossless=0, interleave_mode=0, color_transformation=0, per_frame=True):
    """Encodes n frames (and decodes them when decoded_ptr is not 0).  Returns (seconds, [stream sizes])."""
    sizes = (C.c_size_t * n)()
    job = Job(os.fsencode(library_path), frames_ptr, frame_bytes, streams_ptr, stream_capacity, decoded_ptr or None, sizes, n, threads,
              width, height, bits_per_sample, component_count, near_lossless, interleave_mode, color_transformation,
              1 if per_frame else 0, 0.0)
    errc = _driver().abi_driver_run(C.byref(job))
    if errc != 0:
        raise RuntimeError(f"abi_driver_run: error {errc} from {library_path}")
    return job.seconds, list(sizes)
