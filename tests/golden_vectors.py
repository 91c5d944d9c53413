"""Loaders for the committed golden vectors (made from the unmodified reference by tools/make_golden.py and
tools/make_fixture_golden.py)."""
import ast
import os
from dataclasses import dataclass

import numpy as np

from tests.support import GOLDEN_DIR


@dataclass
class Vector:
    name: str
    bits: int
    near: int
    ilv: int
    xf: int
    pc: tuple
    image: np.ndarray
    ri0: bytes   # the reference's own encoding (no restart markers)
    dec0: np.ndarray
    ri1: bytes   # restart-interval-1 stream stitched from the reference's per-row encodings
    dec1: np.ndarray


def load_vectors():
    z = np.load(os.path.join(GOLDEN_DIR, "reference_vectors.npz"))
    out = []
    for i, m in enumerate(z["meta"]):
        name, bits, near, ilv, xf, pc = ast.literal_eval(str(m))
        out.append(Vector(name, bits, near, ilv, xf, pc, z[f"img_{i}"], z[f"ri0_{i}"].tobytes(), z[f"dec0_{i}"],
                          z[f"ri1_{i}"].tobytes(), z[f"dec1_{i}"]))
    return out


def load_fixture_streams():
    """[(name, stream bytes, reference errc, sha256 of the reference's decoded pixels or None, shape)]"""
    z = np.load(os.path.join(GOLDEN_DIR, "reference_fixture_streams.npz"))
    out = []
    for i, m in enumerate(z["meta"]):
        name, errc, digest, shape = ast.literal_eval(str(m))
        out.append((name, z[f"stream_{i}"].tobytes(), errc, digest, shape))
    return out
