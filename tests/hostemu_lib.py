"""ctypes wrapper of tests/hostemu/libhostemu.so: the product's per-interval device code compiled for the host."""
import ctypes as C
import os
import subprocess

import numpy as np

from tests.support import ROOT

HOSTEMU_DIR = os.path.join(ROOT, "tests", "hostemu")
HOSTEMU_LIB = os.path.join(HOSTEMU_DIR, "libhostemu.so")
CSRC = os.path.join(ROOT, "charls_b200", "csrc")


class HostEmu:
    def __init__(self):
        src = os.path.join(HOSTEMU_DIR, "hostemu.cpp")
        deps = [src] + [os.path.join(CSRC, f) for f in ("jls_codec.cuh", "jls_fast.cuh", "jls_interval.cuh", "jls_common.h", "jls_params.hpp")]
        if not os.path.exists(HOSTEMU_LIB) or any(os.path.getmtime(d) > os.path.getmtime(HOSTEMU_LIB) for d in deps):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I" + CSRC, "-o", HOSTEMU_LIB, src])
        self.dll = C.CDLL(HOSTEMU_LIB)
        self.dll.hostemu_sizeof_params.restype = C.c_size_t
        self.psize = self.dll.hostemu_sizeof_params()
        self.dll.hostemu_encode_scan.restype = C.c_int64
        self.dll.hostemu_encode_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        self.dll.hostemu_decode_scan.restype = C.c_int64
        self.dll.hostemu_decode_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]

    def params(self, sp):
        buf = (C.c_uint8 * self.psize)()
        self.dll.hostemu_make_params(buf, sp.width, sp.height, sp.bits_per_sample, sp.component_count, sp.near_lossless,
                                     sp.interleave_mode, sp.color_transformation, sp.threshold1, sp.threshold2, sp.threshold3,
                                     sp.reset_value, C.c_uint32(sp.restart_interval))
        return buf

    def fast(self, hp):
        return bool(self.dll.hostemu_uses_fast_path(hp))

    def encode(self, hp, plane, capacity, force_general=False):
        plane = np.ascontiguousarray(plane)
        out = np.zeros(capacity, np.uint8)
        n = self.dll.hostemu_encode_scan(hp, plane.ctypes.data, plane.nbytes // plane.shape[0], out.ctypes.data, capacity, int(force_general))
        return n, out[: max(n, 0)].tobytes()

    def decode(self, hp, stream, out, force_general=False):
        src = np.frombuffer(stream, np.uint8).copy()
        return self.dll.hostemu_decode_scan(hp, src.ctypes.data, src.nbytes, out.ctypes.data, out.nbytes // out.shape[0], int(force_general))
