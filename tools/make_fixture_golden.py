"""Packs the reference's restart-marker / conformance streams and the SHA-256 of the reference's decoding of them into
tests/golden/reference_fixture_streams.npz, so that the GPU box (which has no /root/reference) can decode the very
streams the reference's own test-suite uses (test/compliance_test.cpp:43-141) and compare with the reference's output.

Run in the build container:  python tools/make_fixture_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from charls_b200 import codec  # noqa: E402
from tests.support import GOLDEN_DIR, REFERENCE_DATA, reference_library  # noqa: E402

ref = reference_library()
names = [
    "test8_ilv_none_rm_7.jls", "test8_ilv_line_rm_7.jls", "test8_ilv_sample_rm_7.jls", "test8_ilv_sample_rm_300.jls",
    "test16_rm_5.jls", "conformance/t8c0e0.jls", "conformance/t8c1e3.jls", "conformance/t8c2e0.jls", "conformance/t8nde3.jls",
    "conformance/t16e0.jls", "conformance/t16e3.jls", "8bit-monochrome-2x2.jls", "banny-hp3.jls",
    # streams the reference rejects: error-code parity
    "fuzzy-input-bad-run-mode-golomb-code.jls", "fuzzy-input-no-valid-bits-at-the-end.jls",
    "fuzzy_input_golomb_16.jls", "no_start_byte_after_encoded_scan.jls", "conformance/t8sse0.jls",
]
out = {}
meta = []
for i, name in enumerate(names):
    data = open(os.path.join(REFERENCE_DATA, name), "rb").read()
    out[f"stream_{i}"] = np.frombuffer(data, np.uint8)
    try:
        px, fi, ilv = codec.decode(data, lib=ref)
        digest = hashlib.sha256(np.ascontiguousarray(px).tobytes()).hexdigest()
        shape = px.shape
        errc = 0
    except Exception as e:  # CharlsError
        errc = e.errc
    meta.append(repr((name, errc, digest if errc == 0 else None, shape if errc == 0 else None)))
out["meta"] = np.array(meta)
path = os.path.join(GOLDEN_DIR, "reference_fixture_streams.npz")
np.savez_compressed(path, **out)
print(len(names), "->", path, os.path.getsize(path))
