"""C ABI contract on the CPU: the library loads, exports everything include/charls_b200.h declares, keeps the
reference's state machines / error codes, and parses / writes JPEG-LS headers exactly like the reference.  Nothing here
runs a kernel; compute calls are only checked to fail loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from charls_b200 import capi, codec
from charls_b200.capi import CharlsError, FrameInfo, PcParameters, SpiffHeader
from tests import jlsio
from tests.golden_vectors import load_fixture_streams
from tests.support import ROOT, s_mixed


def have_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def test_exports_every_declared_symbol(product):
    header = open(os.path.join(ROOT, "include", "charls_b200.h")).read()
    declared = set(re.findall(r"\b(charlsx?_[a-z0-9_]+)\s*\(", header))
    declared -= {"charls_at_comment_handler", "charls_at_application_data_handler"}
    assert len([d for d in declared if not d.startswith("charlsx_")]) == 48
    for name in sorted(declared):
        assert hasattr(product.dll, name), name
    assert set(capi.ABI_SYMBOLS) <= declared and set(capi.EXT_SYMBOLS) <= declared


def test_struct_sizes():
    # reference include/charls/public_types.h:1075-1078 static_asserts
    assert C.sizeof(SpiffHeader) == 40 and C.sizeof(FrameInfo) == 16 and C.sizeof(PcParameters) == 20
    assert C.sizeof(capi.MappingTableInfo) == 12


def test_version_and_messages(product, reference):
    assert product.version_string() == reference.version_string()
    for errc in list(range(0, 39)) + list(range(100, 113)) + [77]:
        assert product.charls_get_error_message(errc) == reference.charls_get_error_message(errc), errc
    assert product.charls_get_jpegls_category() is not None


def test_null_arguments_and_call_order(product):
    lib = product
    assert lib.charls_jpegls_encoder_set_near_lossless(None, 0) == 101
    assert lib.charls_jpegls_decoder_read_header(None) == 101
    lib.charls_jpegls_encoder_destroy(None)
    lib.charls_jpegls_decoder_destroy(None)
    e = lib.charls_jpegls_encoder_create()
    n = C.c_size_t()
    assert lib.charls_jpegls_encoder_get_estimated_destination_size(e, C.byref(n)) == 100  # frame info first
    assert lib.charls_jpegls_encoder_set_frame_info(e, None) == 101
    for bad, errc in ((FrameInfo(0, 1, 8, 1), 102), (FrameInfo(1, 100001, 8, 1), 103), (FrameInfo(1, 1, 1, 1), 104), (FrameInfo(1, 1, 8, 256), 105)):
        assert lib.charls_jpegls_encoder_set_frame_info(e, C.byref(bad)) == errc
    assert lib.charls_jpegls_encoder_set_interleave_mode(e, 3) == 106
    assert lib.charls_jpegls_encoder_set_near_lossless(e, 256) == 107
    assert lib.charls_jpegls_encoder_set_color_transformation(e, 4) == 109
    assert lib.charls_jpegls_encoder_set_encoding_options(e, 8) == 112
    fi = FrameInfo(4, 2, 8, 1)
    assert lib.charls_jpegls_encoder_set_frame_info(e, C.byref(fi)) == 0
    src = np.zeros(8, np.uint8)
    assert lib.charls_jpegls_encoder_encode_from_buffer(e, src.ctypes.data, 8, 0) == 100  # no destination yet
    assert lib.charls_jpegls_encoder_write_spiff_entry(e, 5, src.ctypes.data, 4) == 100
    dst = np.zeros(256, np.uint8)
    assert lib.charls_jpegls_encoder_set_destination_buffer(e, dst.ctypes.data, 256) == 0
    assert lib.charls_jpegls_encoder_encode_from_buffer(e, src.ctypes.data, 7, 0) == 110  # source too small
    assert lib.charls_jpegls_encoder_encode_from_buffer(e, src.ctypes.data, 8, 3) == 111  # stride too small
    assert lib.charls_jpegls_encoder_rewind(e) == 0
    lib.charls_jpegls_encoder_destroy(e)

    d = lib.charls_jpegls_decoder_create()
    assert lib.charls_jpegls_decoder_read_header(d) == 100
    info = FrameInfo()
    assert lib.charls_jpegls_decoder_get_frame_info(d, C.byref(info)) == 100
    lib.charls_jpegls_decoder_destroy(d)


def test_estimated_destination_size(product, reference):
    """Never below the reference's bound (reference test/jpegls_encoder_test.cpp:278-330 assert lower bounds only)."""
    for w, h, bits, cc in ((1, 1, 2, 1), (100, 100, 8, 1), (4096, 4096, 8, 1), (2048, 2048, 16, 3), (1, 50000, 16, 4)):
        ours, theirs = codec.JpegLSEncoder(product), codec.JpegLSEncoder(reference)
        ours.frame_info(w, h, bits, cc)
        theirs.frame_info(w, h, bits, cc)
        assert ours.estimated_destination_size() >= theirs.estimated_destination_size()
        ours.restart_interval(0)
        assert ours.estimated_destination_size() == theirs.estimated_destination_size()


def test_estimated_destination_size_covers_narrow_tall_images(product, oracle):
    """With one restart interval per line the first sample of every line is predicted from 0 and can take a whole
    LIMIT-bit code word: tall narrow images that the reference squeezes into a few bytes cost 6-10 bytes per line here.
    The estimate must cover them (sizes from the oracle at restart interval 1, no GPU needed)."""
    import numpy as np

    for w, h, bits, cc, value in ((1, 10000, 8, 1, 200), (2, 10000, 16, 1, 51234), (1, 3000, 16, 3, 65535), (4, 5000, 12, 1, 4095),
                                  (3, 2000, 8, 4, 255), (1, 1, 8, 1, 255)):
        dtype = np.uint8 if bits <= 8 else np.uint16
        image = np.full((h, w) if cc == 1 else (h, w, cc), value, dtype)
        ilv = 0 if cc == 1 else 2
        enc = codec.JpegLSEncoder(product)
        enc.frame_info(w, h, bits, cc).interleave_mode(ilv)
        want = oracle.encode_image(image, bits, ilv=ilv, ri=1)
        assert len(want) <= enc.estimated_destination_size(), (w, h, bits, cc, len(want), enc.estimated_destination_size())


def test_offset_table_segments_are_application_data_to_the_reference(product, reference, oracle):
    """The side table of interval offsets (APP11 "JLS-OFFT", extension): both libraries read the same header from a stream
    that carries one, both report the segment to an application-data handler, and the reference decodes the stream to the
    same samples as the stream without it (no GPU needed for any of this)."""
    import numpy as np

    from tests.support import s_mixed

    for img, bits, ri in ((s_mixed(40, 64, 8, seed=3), 8, 1), (s_mixed(23, 31, 12, seed=4), 12, 4)):
        plain = oracle.encode_image(img, bits, ri=ri)
        tabled = jlsio.with_offset_table(plain)
        assert jlsio.without_offset_table(tabled) == plain and len(tabled) > len(plain)
        assert header_summary(product, tabled) == header_summary(reference, tabled) == header_summary(reference, plain)
        want, _, _ = codec.decode(plain, lib=reference)
        got, _, _ = codec.decode(tabled, lib=reference)
        assert np.array_equal(got, want)
        for lib in (product, reference):
            seen = []
            with codec.JpegLSDecoder(lib) as dec:
                dec.at_application_data(lambda app_id, data, size, ctx: seen.append((app_id, C.string_at(data, 8))) or 0)
                dec.source(tabled).read_header()
            assert seen == [(11, b"JLS-OFFT")]


def header_summary(lib, stream):
    """Everything read_header exposes, or the error code."""
    try:
        with codec.JpegLSDecoder(lib) as dec:
            dec.source(stream)
            spiff = dec.read_spiff_header()
            dec.read_header()
            fi = dec.frame_info()
            pc = dec.preset_coding_parameters()
            out = [fi.width, fi.height, fi.bits_per_sample, fi.component_count, dec.color_transformation(), dec.destination_size(),
                   pc.maximum_sample_value, pc.threshold1, pc.threshold2, pc.threshold3, pc.reset_value]
            out += [(dec.near_lossless(c), dec.interleave_mode(c)) for c in range(min(fi.component_count, 4))]
            if spiff is not None:
                out.append(tuple(getattr(spiff, f[0]) for f in SpiffHeader._fields_))
            return out
    except CharlsError as e:
        return e.errc


def make_streams(oracle):
    img = s_mixed(6, 9, 8)
    base = oracle.encode_image(img, 8, ri=0)
    rgb = s_mixed(6, 9, 8, 3, layout="interleaved")
    out = {
        "plain": base,
        "ri": oracle.encode_image(img, 8, ri=2),
        "rgb_line_hp1": oracle.encode_image(rgb, 8, ilv=1, xform=1),
        "rgb_none": oracle.encode_image(s_mixed(6, 9, 8, 3), 8, ilv=0),
        "pc": oracle.encode_image(img, 8, pc=(255, 9, 9, 9, 31)),
        "near": oracle.encode_image(img, 8, near=3),
    }
    seg = lambda m, p: bytes([0xFF, m]) + (len(p) + 2).to_bytes(2, "big") + p
    soi, rest = base[:2], base[2:]
    out["comment_app"] = soi + seg(0xFE, b"hello") + seg(0xE3, b"\x01\x02") + rest
    out["dri3"] = soi + seg(0xDD, b"\x00\x00\x05") + rest
    out["dri4"] = soi + seg(0xDD, b"\x00\x00\x00\x07") + rest
    out["dri_bad"] = soi + seg(0xDD, b"\x05") + rest
    out["two_soi"] = soi + soi + rest
    out["no_soi"] = rest
    out["truncated"] = base[:12]
    out["unknown_marker"] = soi + seg(0x01, b"") + rest
    out["sof_jpeg"] = soi + seg(0xC0, b"\x08\x00\x01\x00\x01\x01\x01\x11\x00") + rest
    out["rst_outside"] = soi + bytes([0xFF, 0xD0]) + rest
    out["lse_ext"] = soi + seg(0xF8, b"\x05") + rest
    out["lse_bad"] = soi + seg(0xF8, b"\x20") + rest
    out["lse_oversize"] = soi + seg(0xF8, b"\x04\x02\x00\x06\x00\x09") + rest
    out["mrfx_bad"] = soi + seg(0xE8, b"mrfx\x09") + rest
    out["mrfx_unsupported"] = soi + seg(0xE8, b"mrfx\x04") + rest
    out["mapping_table"] = soi + seg(0xF8, b"\x02\x01\x01\x00\x01\x02") + seg(0xF8, b"\x03\x01\x01\x03") + rest
    out["mapping_continuation_orphan"] = soi + seg(0xF8, b"\x03\x07\x01\x03") + rest
    out["fill_bytes"] = soi + b"\xff\xff" + rest[0:]
    out["bad_segment_size"] = soi + bytes([0xFF, 0xFE, 0x00, 0x01]) + rest
    spiff = b"SPIFF\x00\x02\x00" + bytes([0, 1]) + (6).to_bytes(4, "big") + (9).to_bytes(4, "big") + bytes([8, 8, 6, 1]) + (96).to_bytes(4, "big") * 2
    eod = (1).to_bytes(4, "big") + b"\xff\xd8"
    out["spiff"] = soi + seg(0xE8, spiff) + seg(0xE8, eod) + rest
    out["spiff_no_eod"] = soi + seg(0xE8, spiff) + rest
    return out


def test_header_parsing_matches_reference(oracle, product, reference):
    """read_spiff_header / read_header: same values, same error codes as the reference (host code only)."""
    streams = make_streams(oracle)
    for name, stream, _, _, _ in load_fixture_streams():
        streams["fixture:" + name] = stream
    for name, stream in streams.items():
        assert header_summary(product, stream) == header_summary(reference, stream), name


def test_callbacks(oracle, product):
    streams = make_streams(oracle)
    seen = []

    def on_comment(data, size, ctx):
        seen.append(("com", C.string_at(data, size)))
        return 0

    def on_app(app_id, data, size, ctx):
        seen.append((app_id, C.string_at(data, size)))
        return 0

    with codec.JpegLSDecoder(product) as dec:
        dec.at_comment(on_comment).at_application_data(on_app)
        dec.source(streams["comment_app"]).read_header()
    assert seen == [("com", b"hello"), (3, b"\x01\x02")]
    with codec.JpegLSDecoder(product) as dec:
        dec.at_comment(lambda d, s, c: 1)
        dec.source(streams["comment_app"])
        with pytest.raises(CharlsError) as info:
            dec.read_header()
        assert info.value.errc == 2


def encoder_header_bytes(lib, configure, image, restart_interval=None):
    """Runs an encode; returns the destination bytes in front of the first entropy-coded byte, whatever the outcome."""
    enc = codec.JpegLSEncoder(lib)
    configure(enc)
    if restart_interval is not None:
        enc.restart_interval(restart_interval)
    dst = np.zeros(4096, np.uint8)
    enc.destination(dst)
    try:
        enc.encode(image)
    except CharlsError as e:
        assert e.errc == 200, e  # only the missing GPU may stop us
    raw = dst.tobytes()
    end = raw.index(b"\xff\xda")
    length = int.from_bytes(raw[end + 2 : end + 4], "big")
    return raw[: end + 2 + length]


def test_encoder_writes_the_reference_header(product, reference):
    """With restart interval 0 the bytes up to the first scan are identical to the reference's; the default (1) only
    adds the DRI segment FF DD 00 04 00 01 in front of SOS (SURVEY.md 8b 'required deviations')."""
    img = s_mixed(5, 7, 8)
    rgb = s_mixed(5, 7, 8, 3, layout="interleaved")
    setups = [
        (lambda e: e.frame_info(7, 5, 8, 1), img),
        (lambda e: e.frame_info(7, 5, 8, 1).near_lossless(2).preset_coding_parameters(255, 9, 9, 9, 31), img),
        (lambda e: e.frame_info(7, 5, 8, 3).interleave_mode(2).color_transformation(1), rgb),
        (lambda e: e.frame_info(7, 5, 8, 3).interleave_mode(1).encoding_options(2), rgb),
        (lambda e: e.frame_info(7, 5, 8, 1).destination(np.zeros(4096, np.uint8)).write_standard_spiff_header(8), img),
        (lambda e: e.frame_info(7, 5, 8, 1).destination(np.zeros(4096, np.uint8)).write_comment(b"abc").write_application_data(2, b"xy"), img),
    ]
    for i, (configure, image) in enumerate(setups[:4]):
        want = encoder_header_bytes(reference, configure, image)
        assert encoder_header_bytes(product, configure, image, 0) == want, i
        with_dri = encoder_header_bytes(product, configure, image)
        sos = want.rindex(b"\xff\xda")
        assert with_dri == want[:sos] + b"\xff\xdd\x00\x04\x00\x01" + want[sos:], i


@pytest.mark.skipif(have_gpu(), reason="only meaningful on a GPU-less host")
def test_compute_fails_loudly_without_gpu(oracle, product):
    """No CPU fallback: the scan codec is CUDA only."""
    with pytest.raises(CharlsError) as info:
        codec.encode(s_mixed(4, 4, 8), 8, lib=product)
    assert info.value.errc == 200
    with pytest.raises(CharlsError) as info:
        codec.decode(oracle.encode_image(s_mixed(4, 4, 8), 8), lib=product)
    assert info.value.errc == 200
    # the two-part forms fail in the first half
    enc = codec.JpegLSEncoder(product)
    enc.frame_info(4, 4, 8, 1)
    enc.destination(np.zeros(4096, np.uint8))
    with pytest.raises(CharlsError) as info:
        enc.encode_begin(s_mixed(4, 4, 8))
    assert info.value.errc == 200
