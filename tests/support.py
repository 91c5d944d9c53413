"""Shared test infrastructure: oracle + reference loaders, synthetic images, stream helpers."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from charls_b200.capi import CharlsLibrary  # noqa: E402
from tests import jlsio  # noqa: E402

ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "libjls_oracle.so")
REF_LIB = os.path.join(ORACLE_DIR, "_ref", "libcharls_ref.so")
REFERENCE_ROOT = "/root/reference"
REFERENCE_DATA = os.path.join(REFERENCE_ROOT, "test", "data")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


class ScanParams(C.Structure):
    _fields_ = [
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("bits_per_sample", C.c_int32),
        ("component_count", C.c_int32),
        ("near_lossless", C.c_int32),
        ("interleave_mode", C.c_int32),
        ("color_transformation", C.c_int32),
        ("threshold1", C.c_int32),
        ("threshold2", C.c_int32),
        ("threshold3", C.c_int32),
        ("reset_value", C.c_int32),
        ("restart_interval", C.c_uint32),
    ]


class Oracle:
    """ctypes wrapper of oracle/libjls_oracle.so (plain-C restatement of the scan codec)."""

    def __init__(self):
        src = os.path.join(ORACLE_DIR, "jls_oracle.c")
        if not os.path.exists(ORACLE_LIB) or os.path.getmtime(ORACLE_LIB) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "libjls_oracle.so"], stdout=subprocess.DEVNULL)
        self.dll = C.CDLL(ORACLE_LIB)
        self.dll.jls_oracle_encode_scan.restype = C.c_int64
        self.dll.jls_oracle_encode_scan.argtypes = [C.POINTER(ScanParams), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        self.dll.jls_oracle_decode_scan.restype = C.c_int64
        self.dll.jls_oracle_decode_scan.argtypes = [C.POINTER(ScanParams), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        self.dll.jls_oracle_default_pc_parameters.restype = None
        self.dll.jls_oracle_default_pc_parameters.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_int32 * 5)]

    def default_pc(self, maxval, near):
        out = (C.c_int32 * 5)()
        self.dll.jls_oracle_default_pc_parameters(maxval, near, C.byref(out))
        return tuple(out)

    def params(self, width, height, bits, ncomp_scan, near=0, ilv=0, xform=0, pc=None, ri=0):
        d = self.default_pc((1 << bits) - 1, near)
        if pc is not None:
            # zero entries mean "default" (reference jpegls_preset_coding_parameters.hpp:88-130)
            maxval = pc[0] if pc[0] else (1 << bits) - 1
            d = self.default_pc(maxval, near)
            d = tuple(p if p else q for p, q in zip(pc, d))
        return ScanParams(width, height, bits, ncomp_scan, near, ilv, xform, d[1], d[2], d[3], d[4], ri)

    def encode_scan(self, p: ScanParams, plane: np.ndarray, stride: int | None = None) -> bytes:
        plane = np.ascontiguousarray(plane)
        if stride is None:
            stride = plane.nbytes // p.height
        cap = plane.nbytes * 5 + 4 * p.height + 1024
        dst = np.empty(cap, dtype=np.uint8)
        n = self.dll.jls_oracle_encode_scan(C.byref(p), plane.ctypes.data, stride, dst.ctypes.data, cap)
        if n < 0:
            raise RuntimeError(f"oracle encode error {n}")
        return dst[:n].tobytes()

    def decode_scan(self, p: ScanParams, data: bytes, out: np.ndarray, stride: int | None = None) -> int:
        src = np.frombuffer(data, dtype=np.uint8)
        if stride is None:
            stride = out.nbytes // p.height
        return self.dll.jls_oracle_decode_scan(C.byref(p), src.ctypes.data, src.nbytes, out.ctypes.data, stride)

    # -- whole-stream helpers built on jlsio ------------------------------------------------------------------
    def encode_image(self, image: np.ndarray, bits: int, *, near=0, ilv=0, xform=0, pc=None, ri=0) -> bytes:
        """image: [H,W] | [C,H,W] (ilv 0) | [H,W,C] (ilv 1/2).  Returns a complete JPEG-LS stream."""
        if image.ndim == 2:
            h, w = image.shape
            ncomp = 1
        elif ilv == 0:
            ncomp, h, w = image.shape
        else:
            h, w, ncomp = image.shape
        scans = []
        if ilv == 0:
            planes = image.reshape(ncomp, h, w)
            for c in range(ncomp):
                p = self.params(w, h, bits, 1, near, 0, 0, pc, ri)
                scans.append((1, near, 0, self.encode_scan(p, planes[c])))
        else:
            p = self.params(w, h, bits, ncomp, near, ilv, xform, pc, ri)
            scans.append((ncomp, near, ilv, self.encode_scan(p, image)))
        pc_seg = None
        if pc is not None:
            pp = self.params(w, h, bits, 1, near, 0, 0, pc, 0)
            pc_seg = (pc[0] if pc[0] else (1 << bits) - 1, pp.threshold1, pp.threshold2, pp.threshold3, pp.reset_value)
        return jlsio.write_stream(w, h, bits, ncomp, scans, color_transformation=xform, pc=pc_seg, restart_interval=ri)

    def decode_image(self, stream: bytes):
        """Returns (pixels, parsed Stream); pixels shaped [H,W] | [C,H,W] | [H,W,C]."""
        s = jlsio.parse(stream)
        dtype = np.uint8 if s.bits_per_sample <= 8 else np.dtype("<u2")
        first = s.scans[0]
        if s.component_count == 1:
            out = np.zeros((s.height, s.width), dtype=dtype)
        elif first.interleave_mode == 0:
            out = np.zeros((s.component_count, s.height, s.width), dtype=dtype)
        else:
            out = np.zeros((s.height, s.width, s.component_count), dtype=dtype)
        comp = 0
        for sc in s.scans:
            p = self.params(
                s.width, s.height, s.bits_per_sample, sc.component_count, sc.near_lossless, sc.interleave_mode,
                s.color_transformation if sc.interleave_mode != 0 else 0, s.pc, sc.restart_interval,
            )
            target = out if (s.component_count == 1 or sc.interleave_mode != 0) else out[comp]
            n = self.decode_scan(p, stream[sc.data_offset :], target)
            if n < 0:
                raise RuntimeError(f"oracle decode error {n}")
            assert sc.data_offset + n == sc.data_end, (sc.data_offset + n, sc.data_end)
            comp += sc.component_count
        return out, s


_oracle = None
_ref = None


def oracle() -> Oracle:
    global _oracle
    if _oracle is None:
        _oracle = Oracle()
    return _oracle


def have_reference_build() -> bool:
    return os.path.exists(REF_LIB)


def reference_library() -> CharlsLibrary:
    """The UNMODIFIED reference, prebuilt by `make -C oracle ref` (travels to the GPU box as a binary)."""
    global _ref
    if _ref is None:
        if not have_reference_build() and os.path.isdir(REFERENCE_ROOT):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "ref"], stdout=subprocess.DEVNULL)
        _ref = CharlsLibrary(REF_LIB, extensions=False)
    return _ref


# -- synthetic images -------------------------------------------------------------------


def s_smooth(h, w, bits, cc=1, seed=1234, layout="planar"):
    """S_smooth: base = 0.8 MAX (0.5 + 0.25 sin(x/97) + 0.25 cos(y/131)); comp c = clip(base(1-0.1c) + N(0, 0.01 MAX))."""
    mx = (1 << bits) - 1
    rng = np.random.default_rng(seed)
    x = np.arange(w, dtype=np.float64)[None, :]
    y = np.arange(h, dtype=np.float64)[:, None]
    base = 0.8 * mx * (0.5 + 0.25 * np.sin(x / 97.0) + 0.25 * np.cos(y / 131.0))
    dtype = np.uint8 if bits <= 8 else np.dtype("<u2")
    comps = [np.clip(base * (1 - 0.1 * c) + rng.normal(0.0, 0.01 * mx, size=(h, w)), 0, mx).astype(dtype) for c in range(cc)]
    if cc == 1:
        return comps[0]
    return np.stack(comps, axis=0) if layout == "planar" else np.stack(comps, axis=-1)


def s_noise(h, w, bits, cc=1, seed=21344, layout="planar"):
    mx = (1 << bits) - 1
    rng = np.random.RandomState(seed)  # mt19937, like the reference's noise tests (test/encode_test.cpp:560-584)
    dtype = np.uint8 if bits <= 8 else np.dtype("<u2")
    shape = (h, w) if cc == 1 else ((cc, h, w) if layout == "planar" else (h, w, cc))
    return rng.randint(0, mx + 1, size=shape).astype(dtype)


def s_mixed(h, w, bits, cc=1, seed=7, layout="planar"):
    """Stress image: flat zero areas (long runs), flat non-zero areas, ramps, noise bursts and saturated pixels."""
    mx = (1 << bits) - 1
    rng = np.random.default_rng(seed)
    dtype = np.uint8 if bits <= 8 else np.dtype("<u2")

    def one(k):
        img = np.zeros((h, w), dtype=np.int64)
        yy, xx = np.mgrid[0:h, 0:w]
        img += ((xx * 3 + yy * 5 + 11 * k) % (mx + 1)) * ((yy // max(1, h // 6)) % 3 == 1)
        img += (mx // 3) * ((yy // max(1, h // 6)) % 3 == 2)
        noise = rng.integers(0, mx + 1, size=(h, w))
        mask = rng.random((h, w)) < 0.08
        img = np.where(mask, noise, img)
        img[:, : w // 5] = np.where(rng.random((h, w // 5)) < 0.02, mx, 0)
        if w > 8:
            img[:, -3:] = mx
        return np.clip(img, 0, mx).astype(dtype)

    comps = [one(k) for k in range(cc)]
    if cc == 1:
        return comps[0]
    return np.stack(comps, axis=0) if layout == "planar" else np.stack(comps, axis=-1)


def read_pnm(path):
    """Tiny P5/P6 reader (reference include/support/portable_anymap_file.hpp); returns [H,W] or [H,W,3], big-endian 16-bit -> native."""
    with open(path, "rb") as f:
        data = f.read()
    tokens = []
    pos = 0
    while len(tokens) < 4:
        while data[pos : pos + 1].isspace():
            pos += 1
        if data[pos : pos + 1] == b"#":
            pos = data.index(b"\n", pos) + 1
            continue
        end = pos
        while not data[end : end + 1].isspace():
            end += 1
        tokens.append(data[pos:end])
        pos = end
    pos += 1
    magic, w, h, mx = tokens[0], int(tokens[1]), int(tokens[2]), int(tokens[3])
    cc = 3 if magic == b"P6" else 1
    if mx > 255:
        arr = np.frombuffer(data, dtype=">u2", count=w * h * cc, offset=pos).astype("<u2")
    else:
        arr = np.frombuffer(data, dtype=np.uint8, count=w * h * cc, offset=pos).copy()
    return (arr.reshape(h, w, cc) if cc == 3 else arr.reshape(h, w)), mx
