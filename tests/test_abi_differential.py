"""Differential test of the host side of the C ABI: random sequences of the calls that run no kernel (parameter setters,
header / table / comment writers, abbreviated format, rewind; on the decoder side header parsing and every getter) are
issued to the B200 library and to the unmodified reference build, and every return code, every byte written and every
value read back must agree.  The one documented difference -- the estimated destination size includes the restart
markers -- is compared through its formula."""
import ctypes as C
import random

import numpy as np
import pytest

from charls_b200.capi import FrameInfo, MappingTableInfo, PcParameters, SpiffHeader
from tests.support import have_reference_build

pytestmark = pytest.mark.skipif(not have_reference_build(), reason="oracle/_ref/libcharls_ref.so not built")


class EncoderPair:
    """The same encoder call on both libraries; asserts equal return codes."""

    def __init__(self, product, reference, capacity=6000):
        self.libs = (product, reference)
        self.handles = [lib.charls_jpegls_encoder_create() for lib in self.libs]
        self.buffers = [np.zeros(capacity, np.uint8) for _ in self.libs]
        self.log = []

    def close(self):
        for lib, handle in zip(self.libs, self.handles):
            lib.charls_jpegls_encoder_destroy(handle)

    def call(self, name, *args, per_library_args=None):
        codes = []
        for i, (lib, handle) in enumerate(zip(self.libs, self.handles)):
            extra = per_library_args[i] if per_library_args else args
            codes.append(getattr(lib, name)(handle, *extra))
        self.log.append((name, args, codes[0]))
        assert codes[0] == codes[1], (name, args, codes, self.log[-12:])
        return codes[0]

    def written(self):
        sizes = []
        for lib, handle in zip(self.libs, self.handles):
            n = C.c_size_t()
            assert lib.charls_jpegls_encoder_get_bytes_written(handle, C.byref(n)) == 0
            sizes.append(n.value)
        assert sizes[0] == sizes[1], (sizes, self.log[-12:])
        assert self.buffers[0][: sizes[0]].tobytes() == self.buffers[1][: sizes[1]].tobytes(), self.log[-12:]
        return self.buffers[0][: sizes[0]].tobytes()


def random_encoder_sequence(pair, rng, steps):
    frame = None
    for _ in range(steps):
        op = rng.randrange(18)
        if op == 0:
            frame = rng.choice([(7, 5, 8, 1), (300, 2, 12, 3), (65535, 1, 16, 4), (1, 1, 2, 1), (9, 9, 8, 200), (0, 1, 8, 1), (1, 1, 17, 1),
                                (1, 1, 8, 0), (100001, 1, 8, 1), (20, 20, 16, 3)])
            info = FrameInfo(*frame)
            if pair.call("charls_jpegls_encoder_set_frame_info", C.byref(info)) != 0:
                frame = None
        elif op == 1:
            pair.call("charls_jpegls_encoder_set_near_lossless", rng.choice([0, 1, 3, 127, 255, 256, -1, 32767]))
        elif op == 2:
            pair.call("charls_jpegls_encoder_set_interleave_mode", rng.choice([-1, 0, 1, 2, 3]))
        elif op == 3:
            pair.call("charls_jpegls_encoder_set_color_transformation", rng.choice([-1, 0, 1, 2, 3, 4]))
        elif op == 4:
            pair.call("charls_jpegls_encoder_set_encoding_options", rng.choice([0, 1, 2, 3, 4, 7, 8, 255]))
        elif op == 5:
            pc = PcParameters(*rng.choice([(0, 0, 0, 0, 0), (255, 3, 7, 21, 64), (255, 9, 9, 9, 31), (4095, 18, 67, 276, 64), (1, 1, 1, 1, 1),
                                           (65535, 1, 2, 3, 255), (255, 22, 7, 21, 64), (70000, 3, 7, 21, 64), (255, 3, 7, 21, 2)]))
            pair.call("charls_jpegls_encoder_set_preset_coding_parameters", C.byref(pc))
        elif op == 6:
            pair.call("charls_jpegls_encoder_set_mapping_table_id", rng.choice([-1, 0, 1, 3, 254, 255]), rng.choice([-1, 0, 1, 255, 256]))
        elif op == 7:
            size = rng.choice([0, 1, 10, 100, len(pair.buffers[0])])
            pair.call("charls_jpegls_encoder_set_destination_buffer", None, None,
                      per_library_args=[(buffer.ctypes.data, size) for buffer in pair.buffers])
        elif op == 8:
            pair.call("charls_jpegls_encoder_write_standard_spiff_header", rng.choice([0, 2, 3, 8, 10, 99]), rng.choice([0, 1, 2, 5]),
                      rng.choice([0, 1, 96]), rng.choice([0, 1, 1024]))
        elif op == 9:
            header = SpiffHeader(rng.choice([0, 1, 5]), rng.choice([1, 3, 4]), rng.choice([0, 5, 20]), rng.choice([0, 7, 300]),
                                 rng.choice([2, 3, 8, 10]), rng.choice([8, 12, 16, 1]), rng.choice([5, 6, 0]), rng.choice([0, 1, 2]), 96, 96)
            pair.call("charls_jpegls_encoder_write_spiff_header", C.byref(header))
        elif op == 10:
            data = bytes(rng.randrange(256) for _ in range(rng.choice([0, 1, 5, 300])))
            pair.call("charls_jpegls_encoder_write_spiff_entry", rng.choice([0, 1, 2, 0xF0, 0xFFFFFFFF]), data, len(data))
        elif op == 11:
            pair.call("charls_jpegls_encoder_write_spiff_end_of_directory_entry")
        elif op == 12:
            data = bytes(rng.randrange(256) for _ in range(rng.choice([0, 1, 13, 70000])))
            pair.call("charls_jpegls_encoder_write_comment", data if data else None, len(data))
        elif op == 13:
            data = bytes(rng.randrange(256) for _ in range(rng.choice([0, 2, 40])))
            pair.call("charls_jpegls_encoder_write_application_data", rng.choice([-1, 0, 7, 15, 16]), data if data else None, len(data))
        elif op == 14:
            entry = rng.choice([0, 1, 2, 3, 255, 256])
            data = bytes(rng.randrange(256) for _ in range(rng.choice([0, 1, 3, 6, 600, 70000])))
            pair.call("charls_jpegls_encoder_write_mapping_table", rng.choice([0, 1, 9, 255, 256]), entry, data if data else None, len(data))
        elif op == 15:
            pair.call("charls_jpegls_encoder_create_abbreviated_format")
        elif op == 16:
            pair.call("charls_jpegls_encoder_rewind")
        else:
            sizes, codes = [], []
            for lib, handle in zip(pair.libs, pair.handles):
                n = C.c_size_t()
                codes.append(lib.charls_jpegls_encoder_get_estimated_destination_size(handle, C.byref(n)))
                sizes.append(n.value)
            assert codes[0] == codes[1], (codes, pair.log[-12:])
            if codes[0] == 0 and frame:
                # the reference's bound plus, per line and component (restart interval 1 by default), a whole LIMIT-bit code
                # word for the first sample, RSTm, padding and a stuffed byte (host/encoder.cpp: estimated_destination_size)
                bits = max(2, frame[2])
                per_interval = (2 * (bits + max(8, bits)) + 7) // 8 + 4
                assert sizes[0] == sizes[1] + per_interval * frame[1] * frame[3] + 8, (sizes, frame)
        pair.written()


@pytest.mark.parametrize("seed", range(12))
def test_random_encoder_call_sequences(product, reference, seed):
    rng = random.Random(1000 + seed)
    for _ in range(25):
        pair = EncoderPair(product, reference)
        try:
            random_encoder_sequence(pair, rng, rng.randrange(4, 40))
        finally:
            pair.close()


def header_streams(reference, rng, count):
    """JPEG-LS streams with assorted marker segments, written by the reference: abbreviated-format table streams and
    complete images with SPIFF headers, comments, application data and mapping tables."""
    streams = []
    for _ in range(count):
        e = reference.charls_jpegls_encoder_create()
        buffer = np.zeros(20000, np.uint8)
        w, h, bits, cc = rng.choice([(6, 4, 8, 1), (5, 3, 8, 3), (4, 4, 12, 1), (3, 2, 16, 3)])
        info = FrameInfo(w, h, bits, cc)
        assert reference.charls_jpegls_encoder_set_frame_info(e, C.byref(info)) == 0
        assert reference.charls_jpegls_encoder_set_destination_buffer(e, buffer.ctypes.data, buffer.size) == 0
        if cc == 3:
            reference.charls_jpegls_encoder_set_interleave_mode(e, rng.choice([0, 1, 2]))
        if rng.random() < 0.4:
            reference.charls_jpegls_encoder_set_near_lossless(e, rng.choice([1, 2]))
        if rng.random() < 0.3:
            pc = PcParameters((1 << bits) - 1, 9, 9, 9, 31)
            reference.charls_jpegls_encoder_set_preset_coding_parameters(e, C.byref(pc))
        if rng.random() < 0.5:
            reference.charls_jpegls_encoder_write_standard_spiff_header(e, 8 if cc == 1 else 10, 1, 96, 96)
            if rng.random() < 0.5:
                reference.charls_jpegls_encoder_write_spiff_entry(e, 7, b"entry", 5)
        if rng.random() < 0.6:
            reference.charls_jpegls_encoder_write_comment(e, b"a comment", 9)
        if rng.random() < 0.5:
            reference.charls_jpegls_encoder_write_application_data(e, rng.randrange(16), b"\x01\x02\x03", 3)
        if rng.random() < 0.5:
            table = bytes(rng.randrange(256) for _ in range(rng.choice([3, 30, 700])))
            reference.charls_jpegls_encoder_write_mapping_table(e, rng.choice([1, 5]), rng.choice([1, 3]), table, len(table))
            if rng.random() < 0.5:
                reference.charls_jpegls_encoder_set_mapping_table_id(e, 0, 1)
        # abbreviated format (tables only) where the reference accepts it in this state, else a complete image
        if not (rng.random() < 0.3 and reference.charls_jpegls_encoder_create_abbreviated_format(e) == 0):
            sample_bytes = 1 if bits <= 8 else 2
            pixels = np.frombuffer(bytes(rng.randrange(256) for _ in range(w * h * cc * sample_bytes)), np.uint8).copy()
            if bits not in (8, 16):
                view = pixels.view(np.uint16)
                view &= (1 << bits) - 1
            rc = reference.charls_jpegls_encoder_encode_from_buffer(e, pixels.ctypes.data, pixels.size, 0)
            assert rc == 0, rc
        n = C.c_size_t()
        reference.charls_jpegls_encoder_get_bytes_written(e, C.byref(n))
        reference.charls_jpegls_encoder_destroy(e)
        streams.append(buffer[: n.value].tobytes())
    return streams


def decoder_observations(lib, stream, truncate=None):
    """Everything the decoder tells about a stream without decoding it."""
    data = np.frombuffer(stream if truncate is None else stream[:truncate], np.uint8).copy()
    d = lib.charls_jpegls_decoder_create()
    seen = []
    try:
        seen.append(lib.charls_jpegls_decoder_read_header(d))  # no source yet: invalid_operation
        seen.append(lib.charls_jpegls_decoder_set_source_buffer(d, data.ctypes.data, data.size))
        spiff, found = SpiffHeader(), C.c_int(-1)
        rc = lib.charls_jpegls_decoder_read_spiff_header(d, C.byref(spiff), C.byref(found))
        seen.append(("spiff", rc, found.value, bytes(spiff) if rc == 0 and found.value else None))
        rc = lib.charls_jpegls_decoder_read_header(d)
        seen.append(("header", rc))
        info = FrameInfo()
        seen.append((lib.charls_jpegls_decoder_get_frame_info(d, C.byref(info)), info.width, info.height, info.bits_per_sample,
                     info.component_count))
        for component in (0, 1, 2, 7):
            value = C.c_int(-9)
            seen.append(("near", lib.charls_jpegls_decoder_get_near_lossless(d, component, C.byref(value)), value.value))
            value = C.c_int(-9)
            seen.append(("ilv", lib.charls_jpegls_decoder_get_interleave_mode(d, component, C.byref(value)), value.value))
            value = C.c_int(-9)
            seen.append(("table id", lib.charls_decoder_get_mapping_table_id(d, component, C.byref(value)), value.value))
        pc = PcParameters()
        seen.append(("pc", lib.charls_jpegls_decoder_get_preset_coding_parameters(d, 0, C.byref(pc)), bytes(pc)))
        value = C.c_int(-9)
        seen.append(("transform", lib.charls_jpegls_decoder_get_color_transformation(d, C.byref(value)), value.value))
        for stride in (0, 64, 3):
            size = C.c_size_t(0)
            seen.append(("size", lib.charls_jpegls_decoder_get_destination_size(d, stride, C.byref(size)), size.value))
        value = C.c_int(-9)
        seen.append(("format", lib.charls_decoder_get_compressed_data_format(d, C.byref(value)), value.value))
        count = C.c_int(-9)
        seen.append(("tables", lib.charls_decoder_get_mapping_table_count(d, C.byref(count)), count.value))
        for table_id in (1, 5, 9):
            index = C.c_int(-9)
            seen.append(("find", lib.charls_decoder_find_mapping_table_index(d, table_id, C.byref(index)), index.value))
        for index in range(max(count.value, 0) + 1):
            table = MappingTableInfo()
            rc = lib.charls_decoder_get_mapping_table_info(d, index, C.byref(table))
            seen.append(("table", rc, bytes(table) if rc == 0 else None))
            if rc == 0:
                out = np.zeros(table.data_size + 4, np.uint8)
                rc = lib.charls_decoder_get_mapping_table_data(d, index, out.ctypes.data, table.data_size)
                seen.append(("table data", rc, out.tobytes()))
                seen.append(("table data, short buffer", lib.charls_decoder_get_mapping_table_data(d, index, out.ctypes.data, 1)))
    finally:
        lib.charls_jpegls_decoder_destroy(d)
    return seen


@pytest.mark.parametrize("seed", range(6))
def test_decoder_header_observations(product, reference, seed):
    rng = random.Random(77 + seed)
    for stream in header_streams(reference, rng, 20):
        assert decoder_observations(product, stream) == decoder_observations(reference, stream)
        # cut anywhere: the same error from the same call
        for _ in range(6):
            cut = rng.randrange(0, len(stream))
            assert decoder_observations(product, stream, cut) == decoder_observations(reference, stream, cut), cut
        # damage one byte of the marker segments
        first_scan = stream.find(b"\xff\xda")
        limit = first_scan if first_scan > 0 else len(stream)
        for _ in range(6):
            damaged = bytearray(stream)
            damaged[rng.randrange(2, limit)] = rng.randrange(256)
            assert decoder_observations(product, bytes(damaged)) == decoder_observations(reference, bytes(damaged))


def test_argument_checks_of_the_compute_calls(product, reference):
    """Sizes and strides are validated before any sample is touched: the calls that must fail fail with the reference's
    error code (those that would succeed need a GPU here and are covered by tests/test_gpu_parity.py)."""
    cases = []
    for (w, h, bits, cc, ilv) in ((8, 4, 8, 1, 0), (8, 4, 8, 3, 0), (8, 4, 8, 3, 1), (8, 4, 8, 3, 2), (5, 3, 16, 1, 0), (5, 3, 12, 4, 2)):
        sample_bytes = 1 if bits <= 8 else 2
        row = w * sample_bytes * (cc if ilv != 0 else 1)
        full = row * h * (cc if ilv == 0 else 1)
        for size, stride in ((0, 0), (1, 0), (full - 1, 0), (full, row - 1), (full, row + 3), (full - 1, row), (full + row, row + 3)):
            cases.append((w, h, bits, cc, ilv, size, stride))
    for (w, h, bits, cc, ilv, size, stride) in cases:
        outcomes = []
        for lib in (product, reference):
            e = lib.charls_jpegls_encoder_create()
            info = FrameInfo(w, h, bits, cc)
            destination = np.zeros(4096, np.uint8)
            source = np.zeros(max(size, 1) + 64, np.uint8)
            assert lib.charls_jpegls_encoder_set_frame_info(e, C.byref(info)) == 0
            assert lib.charls_jpegls_encoder_set_interleave_mode(e, ilv) == 0
            assert lib.charls_jpegls_encoder_set_destination_buffer(e, destination.ctypes.data, destination.size) == 0
            rc = lib.charls_jpegls_encoder_encode_from_buffer(e, source.ctypes.data, size, stride)
            null_rc = lib.charls_jpegls_encoder_encode_from_buffer(e, None, size, stride) if rc != 0 else None
            lib.charls_jpegls_encoder_destroy(e)
            outcomes.append((rc, null_rc))
        if outcomes[1][0] != 0:  # the reference rejects the arguments: same code from us
            assert outcomes[0] == outcomes[1], (w, h, bits, cc, ilv, size, stride, outcomes)
        else:
            assert outcomes[0][0] in (0, 200), (w, h, bits, cc, ilv, size, stride, outcomes)

    # decoder: destination size and stride checks against a reference-made stream
    for (w, h, bits, cc, ilv) in ((8, 4, 8, 1, 0), (8, 4, 8, 3, 2), (5, 3, 16, 3, 0), (5, 3, 16, 3, 1)):
        e = reference.charls_jpegls_encoder_create()
        info = FrameInfo(w, h, bits, cc)
        sample_bytes = 1 if bits <= 8 else 2
        pixels = np.arange(w * h * cc * sample_bytes, dtype=np.uint8)
        buffer = np.zeros(4096, np.uint8)
        reference.charls_jpegls_encoder_set_frame_info(e, C.byref(info))
        reference.charls_jpegls_encoder_set_interleave_mode(e, ilv)
        reference.charls_jpegls_encoder_set_destination_buffer(e, buffer.ctypes.data, buffer.size)
        assert reference.charls_jpegls_encoder_encode_from_buffer(e, pixels.ctypes.data, pixels.size, 0) == 0
        n = C.c_size_t()
        reference.charls_jpegls_encoder_get_bytes_written(e, C.byref(n))
        reference.charls_jpegls_encoder_destroy(e)
        stream = buffer[: n.value].copy()
        row = w * sample_bytes * (cc if ilv != 0 else 1)
        full = pixels.size
        for size, stride in ((0, 0), (full - 1, 0), (full, row - 1), (full, row + 2), (full - 1, row), (full + 64, row + 2)):
            outcomes = []
            for lib in (product, reference):
                d = lib.charls_jpegls_decoder_create()
                out = np.zeros(full + 256, np.uint8)
                assert lib.charls_jpegls_decoder_set_source_buffer(d, stream.ctypes.data, stream.size) == 0
                premature = lib.charls_jpegls_decoder_decode_to_buffer(d, out.ctypes.data, size, stride)  # header not read yet
                assert lib.charls_jpegls_decoder_read_header(d) == 0
                rc = lib.charls_jpegls_decoder_decode_to_buffer(d, out.ctypes.data, size, stride)
                lib.charls_jpegls_decoder_destroy(d)
                outcomes.append((premature, rc))
            assert outcomes[0][0] == outcomes[1][0], (w, h, bits, cc, ilv, size, stride, outcomes)
            if outcomes[1][1] != 0:
                # a multi-scan frame fails at the scan whose plane no longer fits, after the earlier ones were decoded:
                # that needs the GPU (error 200 without one)
                later_scan = ilv == 0 and cc > 1 and outcomes[0][1] == 200 and size >= row * h
                assert later_scan or outcomes[0][1] == outcomes[1][1], (w, h, bits, cc, ilv, size, stride, outcomes)
            else:
                assert outcomes[0][1] in (0, 200), (w, h, bits, cc, ilv, size, stride, outcomes)


def test_validate_spiff_header_matches_reference(product, reference):
    rng = random.Random(5)
    for _ in range(3000):
        header = SpiffHeader(rng.choice([0, 1, 2]), rng.choice([0, 1, 3, 4, 255]), rng.choice([0, 1, 4, 9]), rng.choice([0, 1, 6, 9]),
                             rng.choice([0, 1, 2, 3, 4, 8, 10, 13, 14, 15]), rng.choice([0, 1, 2, 8, 12, 16, 17]),
                             rng.choice([0, 4, 5, 6, 7]), rng.choice([0, 1, 2, 3]), rng.choice([0, 1, 96]), rng.choice([0, 1, 96]))
        info = FrameInfo(rng.choice([1, 6, 9]), rng.choice([1, 4, 9]), rng.choice([2, 8, 12, 16]), rng.choice([1, 3, 4]))
        assert product.charls_validate_spiff_header(C.byref(header), C.byref(info)) == reference.charls_validate_spiff_header(
            C.byref(header), C.byref(info)), bytes(header)
