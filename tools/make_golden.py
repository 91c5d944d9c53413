"""Generates tests/golden/reference_vectors.npz from the UNMODIFIED reference (oracle/_ref/libcharls_ref.so).

Run in the build container (where /root/reference exists):  python tools/make_golden.py
Every vector holds: the source image, the reference's own encoding (no restart markers -- the only thing it can
write), the reference's decoding of it, and a restart-interval-1 stream stitched from the reference's encodings of every
single row as a W x 1 image together with the reference's decoding of that stream.
The stitched stream is byte-for-byte what a restart-interval-1 encoder must produce.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from charls_b200 import codec  # noqa: E402
from tests import jlsio  # noqa: E402
from tests.support import GOLDEN_DIR, reference_library, s_mixed, s_noise, s_smooth  # noqa: E402

ref = reference_library()


def ref_encode(img, bits, near, ilv, xf, pc):
    return codec.encode(img, bits, near_lossless=near, interleave_mode=ilv, color_transformation=xf, preset=pc, lib=ref)


def stitched_ri1(img, bits, near, ilv, xf, pc):
    """Restart-interval-1 stream made of the reference's encodings of the single rows."""
    if img.ndim == 2:
        h, w, cc = img.shape[0], img.shape[1], 1
    elif ilv == 0:
        cc, h, w = img.shape
    else:
        h, w, cc = img.shape
    scans = []
    if ilv == 0:
        planes = img.reshape(cc, h, w)
        for c in range(cc):
            payload = bytearray()
            for r in range(h):
                s = ref_encode(planes[c, r : r + 1], bits, near, 0, 0, pc)
                p = jlsio.parse(s)
                payload += s[p.scans[0].data_offset : p.scans[0].data_end]
                if r + 1 < h:
                    payload += bytes([0xFF, 0xD0 + (r % 8)])
            scans.append((1, near, 0, bytes(payload)))
    else:
        payload = bytearray()
        for r in range(h):
            s = ref_encode(img[r : r + 1], bits, near, ilv, xf, pc)
            p = jlsio.parse(s)
            payload += s[p.scans[0].data_offset : p.scans[0].data_end]
            if r + 1 < h:
                payload += bytes([0xFF, 0xD0 + (r % 8)])
        scans.append((cc, near, ilv, bytes(payload)))
    pc_seg = None
    if pc is not None:
        q = jlsio.parse(ref_encode(img, bits, near, ilv, xf, pc))
        pc_seg = q.pc
    return jlsio.write_stream(w, h, bits, cc, scans, color_transformation=xf, pc=pc_seg, restart_interval=1)


cases = []


def add(name, img, bits, near=0, ilv=0, xf=0, pc=None):
    ri0 = ref_encode(img, bits, near, ilv, xf, pc)
    dec0, _, _ = codec.decode(ri0, lib=ref)
    ri1 = stitched_ri1(img, bits, near, ilv, xf, pc)
    dec1, _, _ = codec.decode(ri1, lib=ref)
    if near == 0:
        mask = (1 << bits) - 1
        assert np.array_equal(dec1, img & mask if xf == 0 else img), name
    cases.append(dict(name=name, bits=bits, near=near, ilv=ilv, xf=xf, pc=pc, image=img, ri0=ri0, dec0=dec0, ri1=ri1, dec1=dec1))


add("mono8_smooth", s_smooth(24, 67, 8), 8)
add("mono8_mixed", s_mixed(24, 67, 8), 8)
add("mono8_noise", s_noise(16, 40, 8), 8)
add("mono8_near3", s_mixed(20, 50, 8), 8, near=3)
add("mono12_near2", s_smooth(20, 61, 12), 12, near=2)
add("mono12_mixed", s_mixed(20, 61, 12), 12)
add("mono16_noise", s_noise(12, 33, 16), 16)
add("mono16_smooth", s_smooth(16, 48, 16), 16)
add("mono2", s_mixed(10, 30, 2), 2)
add("mono5_near1", s_mixed(10, 30, 5), 5, near=1)
add("mono10", s_smooth(12, 40, 10), 10)
add("rgb8_none", s_mixed(12, 31, 8, 3, layout="planar"), 8, ilv=0)
add("rgb8_line", s_mixed(12, 31, 8, 3, layout="interleaved"), 8, ilv=1)
add("rgb8_sample", s_mixed(12, 31, 8, 3, layout="interleaved"), 8, ilv=2)
add("rgb8_sample_near2", s_smooth(12, 31, 8, 3, layout="interleaved"), 8, near=2, ilv=2)
add("rgb8_line_hp1", s_smooth(12, 31, 8, 3, layout="interleaved"), 8, ilv=1, xf=1)
add("rgb8_sample_hp2", s_mixed(12, 31, 8, 3, layout="interleaved"), 8, ilv=2, xf=2)
add("rgb8_sample_hp3", s_mixed(12, 31, 8, 3, layout="interleaved"), 8, ilv=2, xf=3)
add("rgb16_sample_hp1", s_smooth(10, 29, 16, 3, layout="interleaved"), 16, ilv=2, xf=1)
add("rgb16_sample_hp1_noise", s_noise(8, 21, 16, 3, layout="interleaved"), 16, ilv=2, xf=1)
add("two8_sample", s_mixed(9, 20, 8, 2, layout="interleaved"), 8, ilv=2)
add("four8_line", s_mixed(9, 20, 8, 4, layout="interleaved"), 8, ilv=1)
add("four16_sample", s_noise(6, 15, 16, 4, layout="interleaved"), 16, ilv=2)
add("mono8_preset", s_mixed(16, 40, 8), 8, pc=(255, 9, 9, 9, 31))
add("mono8_preset_near", s_mixed(16, 40, 8), 8, near=2, pc=(0, 0, 0, 0, 3))
# the rows of SURVEY.md Appendix B
add("kat_zeros", np.zeros((1, 4), np.uint8), 8)
add("kat_1234", np.array([[1, 2, 3, 4]], np.uint8), 8)
add("kat_255x8", np.full((1, 8), 255, np.uint8), 8)
add("kat_run_then_9", np.array([[0, 0, 0, 0, 0, 0, 0, 9]], np.uint8), 8)
add("kat_ramp_ff_stuffing", np.arange(0, 256, 5, dtype=np.uint8)[None, :], 8)
add("kat_12bit_near2", np.array([[100, 104, 97, 110, 2000, 2003, 1999, 4095]], np.dtype("<u2")), 12, near=2)
add("kat_rgb16_hp1", np.array([[(1000, 2000, 3000), (1010, 2010, 3010), (65535, 0, 32768)]], np.dtype("<u2")), 16, ilv=2, xf=1)

os.makedirs(GOLDEN_DIR, exist_ok=True)
out = {}
meta = []
for i, c in enumerate(cases):
    out[f"img_{i}"] = c["image"]
    out[f"ri0_{i}"] = np.frombuffer(c["ri0"], np.uint8)
    out[f"ri1_{i}"] = np.frombuffer(c["ri1"], np.uint8)
    out[f"dec0_{i}"] = c["dec0"]
    out[f"dec1_{i}"] = c["dec1"]
    meta.append(repr((c["name"], c["bits"], c["near"], c["ilv"], c["xf"], c["pc"])))
out["meta"] = np.array(meta)
path = os.path.join(GOLDEN_DIR, "reference_vectors.npz")
np.savez_compressed(path, **out)
print(len(cases), "vectors ->", path, os.path.getsize(path), "bytes")
for c in cases[-7:]:
    p = jlsio.parse(c["ri1"])
    print(c["name"], c["ri1"][p.scans[0].data_offset : p.scans[0].data_end].hex())
