"""Pins the plain-C oracle (oracle/jls_oracle.c): golden vectors made by the reference, the reference's own fixtures,
and -- where the unmodified reference build is available -- byte-for-byte comparison on fresh inputs."""
import glob
import hashlib
import os

import numpy as np
import pytest

from charls_b200 import codec
from tests import jlsio
from tests.golden_vectors import load_fixture_streams
from tests.support import REFERENCE_DATA, read_pnm, s_mixed, s_noise, s_smooth


def scan_payloads(stream):
    s = jlsio.parse(stream)
    return [stream[sc.data_offset : sc.data_end] for sc in s.scans]


def test_golden_encode_no_restart(oracle, golden):
    """oracle encoder == reference encoder, entropy bytes of every scan (reference compliance_test.cpp:171-196 style)."""
    for v in golden:
        got = oracle.encode_image(v.image, v.bits, near=v.near, ilv=v.ilv, xform=v.xf, pc=v.pc, ri=0)
        assert scan_payloads(got) == scan_payloads(v.ri0), v.name


def test_golden_encode_restart_interval_1(oracle, golden):
    """oracle Ri=1 encoder == streams stitched from the reference's per-row encodings."""
    for v in golden:
        got = oracle.encode_image(v.image, v.bits, near=v.near, ilv=v.ilv, xform=v.xf, pc=v.pc, ri=1)
        assert scan_payloads(got) == scan_payloads(v.ri1), v.name


def test_golden_decode(oracle, golden):
    for v in golden:
        for stream, want in ((v.ri0, v.dec0), (v.ri1, v.dec1)):
            got, _ = oracle.decode_image(stream)
            assert np.array_equal(got, want), v.name


def test_appendix_b_known_answers(golden):
    """SURVEY.md Appendix B rows, taken from the reference at survey time, reproduced by tools/make_golden.py."""
    expected = {
        "kat_zeros": "f0", "kat_1234": "5b68", "kat_255x8": "495e", "kat_run_then_9": "fa14",
        "kat_ramp_ff_stuffing": "8a52fe97f7ff36db6dff7fff7fff7fff7ff0",
        "kat_12bit_near2": "178e6800000000379428000000001c7c",
        "kat_rgb16_hp1": "000000000001f82f8fa01e0be00427010f2c0770000000000003f058",
    }
    by_name = {v.name: v for v in golden}
    for name, hexbytes in expected.items():
        assert scan_payloads(by_name[name].ri1)[0].hex() == hexbytes, name


def test_reference_fixture_streams(oracle):
    """The reference's restart / conformance streams decode to what the reference decodes; corrupt ones are rejected."""
    for name, stream, errc, digest, shape in load_fixture_streams():
        if errc == 0:
            got, _ = oracle.decode_image(stream)
            assert got.shape == tuple(shape), name
            assert hashlib.sha256(np.ascontiguousarray(got).tobytes()).hexdigest() == digest, name
        elif errc in (4, 5):  # the entropy decoder's errors; container-level errors are the product host code's job
            with pytest.raises(RuntimeError):
                oracle.decode_image(stream)


@pytest.mark.skipif(not os.path.isdir(REFERENCE_DATA), reason="reference checkout not present")
def test_reference_fixtures_in_place(oracle, reference):
    """Every decodable .jls under the reference's test/data: oracle == reference; Annex-E images == their .ppm/.pgm."""
    files = sorted(glob.glob(REFERENCE_DATA + "/*.jls") + glob.glob(REFERENCE_DATA + "/conformance/*.jls"))
    decoded = 0
    for path in files:
        data = open(path, "rb").read()
        try:
            want, fi, ilv = codec.decode(data, lib=reference)
        except Exception:
            continue
        got, _ = oracle.decode_image(data)
        assert np.array_equal(got, want), path
        decoded += 1
    assert decoded >= 21
    # lossless Annex-E streams reproduce the source images (reference compliance_test.cpp:43-98)
    ppm, _ = read_pnm(os.path.join(REFERENCE_DATA, "conformance", "test8.ppm"))
    got, _ = oracle.decode_image(open(os.path.join(REFERENCE_DATA, "conformance", "t8c2e0.jls"), "rb").read())
    assert np.array_equal(got, ppm)
    got, _ = oracle.decode_image(open(os.path.join(REFERENCE_DATA, "conformance", "t8c0e0.jls"), "rb").read())
    assert np.array_equal(got, ppm.transpose(2, 0, 1))
    pgm, _ = read_pnm(os.path.join(REFERENCE_DATA, "conformance", "test16.pgm"))
    got, _ = oracle.decode_image(open(os.path.join(REFERENCE_DATA, "conformance", "t16e0.jls"), "rb").read())
    assert np.array_equal(got, pgm)
    # re-encoding is byte-identical to the fixture (reference compliance_test.cpp:171-196)
    enc = oracle.encode_image(ppm, 8, ilv=2)
    fixture = open(os.path.join(REFERENCE_DATA, "conformance", "t8c2e0.jls"), "rb").read()
    assert scan_payloads(enc) == scan_payloads(fixture)


def test_against_reference_build(oracle, reference):
    """Fresh inputs: oracle encoder bytes == reference encoder bytes; reference decodes the oracle's restart streams."""
    cases = []
    for bits in (2, 3, 7, 8, 9, 12, 15, 16):
        for gen in (s_smooth, s_noise, s_mixed):
            cases.append((gen(19, 41, bits), bits, 0, 0, 0))
            if bits >= 3:
                cases.append((gen(19, 41, bits), bits, 1 if bits < 6 else 3, 0, 0))
    for bits in (8, 16):
        for cc in (2, 3, 4):
            for ilv in (0, 1, 2):
                img = s_mixed(13, 29, bits, cc, layout="planar" if ilv == 0 else "interleaved")
                cases.append((img, bits, 0, ilv, 0))
                cases.append((img, bits, 2, ilv, 0))
                if cc == 3 and ilv != 0:
                    cases += [(img, bits, 0, ilv, xf) for xf in (1, 2, 3)]
    for img, bits, near, ilv, xf in cases:
        a = codec.encode(img, bits, near_lossless=near, interleave_mode=ilv, color_transformation=xf, lib=reference)
        b = oracle.encode_image(img, bits, near=near, ilv=ilv, xform=xf)
        assert scan_payloads(a) == scan_payloads(b), (img.shape, bits, near, ilv, xf)
        for ri in (1, 5):
            c = oracle.encode_image(img, bits, near=near, ilv=ilv, xform=xf, ri=ri)
            want, _, _ = codec.decode(c, lib=reference)
            got, _ = oracle.decode_image(c)
            assert np.array_equal(got, want), (img.shape, bits, near, ilv, xf, ri)
            if near == 0:
                assert np.array_equal(want, img & ((1 << bits) - 1) if xf == 0 else img)
            else:
                assert int(np.abs(want.astype(np.int64) - img.astype(np.int64)).max()) <= near
