"""Host-side mirror of the reference's C++ wrapper classes over the C ABI.

``JpegLSEncoder`` / ``JpegLSDecoder`` follow charls::jpegls_encoder / charls::jpegls_decoder
(reference include/charls/jpegls_encoder.hpp:58-446, include/charls/jpegls_decoder.hpp:108-538): same method names,
argument meaning and error behaviour (a non-zero ``charls_jpegls_errc`` raises ``CharlsError``).  Every method is
a thin call into the shared library given by ``lib`` -- by default the in-tree CUDA build.
"""
from __future__ import annotations

import ctypes as C
from ctypes import byref, c_int32, c_size_t, c_uint32

import numpy as np

from .capi import CharlsLibrary, FrameInfo, PcParameters, SpiffHeader, default_library

INTERLEAVE_NONE, INTERLEAVE_LINE, INTERLEAVE_SAMPLE = 0, 1, 2
TRANSFORM_NONE, TRANSFORM_HP1, TRANSFORM_HP2, TRANSFORM_HP3 = 0, 1, 2, 3


def _as_buffer(data):
    """Returns (address, size_in_bytes, keepalive) for bytes / bytearray / numpy arrays."""
    if isinstance(data, np.ndarray):
        if not data.flags["C_CONTIGUOUS"]:
            data = np.ascontiguousarray(data)
        return data.ctypes.data, data.nbytes, data
    if isinstance(data, (bytes, bytearray, memoryview)):
        arr = np.frombuffer(data, dtype=np.uint8)
        return arr.ctypes.data, arr.nbytes, arr
    raise TypeError(f"unsupported buffer type {type(data)!r}")


class JpegLSEncoder:
    def __init__(self, lib: CharlsLibrary | None = None):
        self.lib = lib or default_library()
        self._h = self.lib.charls_jpegls_encoder_create()
        if not self._h:
            raise MemoryError("charls_jpegls_encoder_create failed")
        self._destination = None

    def close(self):
        if self._h:
            self.lib.charls_jpegls_encoder_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- configuration (jpegls_encoder.hpp:92-215)
    def frame_info(self, width, height, bits_per_sample, component_count):
        fi = FrameInfo(width, height, bits_per_sample, component_count)
        self.lib.check(self.lib.charls_jpegls_encoder_set_frame_info(self._h, byref(fi)))
        return self

    def near_lossless(self, near):
        self.lib.check(self.lib.charls_jpegls_encoder_set_near_lossless(self._h, near))
        return self

    def interleave_mode(self, mode):
        self.lib.check(self.lib.charls_jpegls_encoder_set_interleave_mode(self._h, mode))
        return self

    def color_transformation(self, transformation):
        self.lib.check(self.lib.charls_jpegls_encoder_set_color_transformation(self._h, transformation))
        return self

    def encoding_options(self, options):
        self.lib.check(self.lib.charls_jpegls_encoder_set_encoding_options(self._h, options))
        return self

    def preset_coding_parameters(self, maximum_sample_value, t1, t2, t3, reset_value):
        pc = PcParameters(maximum_sample_value, t1, t2, t3, reset_value)
        self.lib.check(self.lib.charls_jpegls_encoder_set_preset_coding_parameters(self._h, byref(pc)))
        return self

    def offset_table(self, enabled=True):
        """Side table of interval offsets in front of every scan (extension, include/charls_b200.h)."""
        self.lib.check(self.lib.charlsx_jpegls_encoder_set_offset_table(self._h, 1 if enabled else 0))
        return self

    def restart_interval(self, lines):
        """B200 extension (charlsx_jpegls_encoder_set_restart_interval); the reference cannot encode restart markers."""
        self.lib.check(self.lib.charlsx_jpegls_encoder_set_restart_interval(self._h, lines))
        return self

    def estimated_destination_size(self) -> int:
        size = c_size_t()
        self.lib.check(self.lib.charls_jpegls_encoder_get_estimated_destination_size(self._h, byref(size)))
        return size.value

    def destination(self, buffer: np.ndarray):
        addr, size, keep = _as_buffer(buffer)
        self.lib.check(self.lib.charls_jpegls_encoder_set_destination_buffer(self._h, addr, size))
        self._destination = keep
        return self

    def write_standard_spiff_header(self, color_space, resolution_units=0, vertical_resolution=1, horizontal_resolution=1):
        self.lib.check(
            self.lib.charls_jpegls_encoder_write_standard_spiff_header(
                self._h, color_space, resolution_units, vertical_resolution, horizontal_resolution
            )
        )
        return self

    def write_spiff_header(self, header: SpiffHeader):
        self.lib.check(self.lib.charls_jpegls_encoder_write_spiff_header(self._h, byref(header)))
        return self

    def write_spiff_entry(self, tag: int, data: bytes):
        buf = C.create_string_buffer(data, len(data))
        self.lib.check(self.lib.charls_jpegls_encoder_write_spiff_entry(self._h, tag, C.addressof(buf), len(data)))
        return self

    def write_spiff_end_of_directory_entry(self):
        self.lib.check(self.lib.charls_jpegls_encoder_write_spiff_end_of_directory_entry(self._h))
        return self

    def write_comment(self, data: bytes):
        buf = C.create_string_buffer(data, len(data)) if data else None
        self.lib.check(self.lib.charls_jpegls_encoder_write_comment(self._h, C.addressof(buf) if buf else None, len(data)))
        return self

    def write_application_data(self, app_id: int, data: bytes):
        buf = C.create_string_buffer(data, len(data)) if data else None
        self.lib.check(
            self.lib.charls_jpegls_encoder_write_application_data(self._h, app_id, C.addressof(buf) if buf else None, len(data))
        )
        return self

    def write_mapping_table(self, table_id: int, entry_size: int, data: bytes):
        buf = C.create_string_buffer(data, len(data))
        self.lib.check(self.lib.charls_jpegls_encoder_write_mapping_table(self._h, table_id, entry_size, C.addressof(buf), len(data)))
        return self

    def set_mapping_table_id(self, component_index: int, table_id: int):
        self.lib.check(self.lib.charls_jpegls_encoder_set_mapping_table_id(self._h, component_index, table_id))
        return self

    # -- coding (jpegls_encoder.hpp:330-410)
    def encode(self, source, stride: int = 0) -> int:
        addr, size, _keep = _as_buffer(source)
        self.lib.check(self.lib.charls_jpegls_encoder_encode_from_buffer(self._h, addr, size, stride))
        return self.bytes_written()

    # two-part form (extension, include/charls_b200.h): several encoder objects in flight from one thread
    def encode_begin(self, source, stride: int = 0):
        addr, size, keep = _as_buffer(source)
        self._source_in_flight = keep
        self.lib.check(self.lib.charlsx_jpegls_encoder_encode_from_buffer_begin(self._h, addr, size, stride))
        return self

    def encode_end(self) -> int:
        self.lib.check(self.lib.charlsx_jpegls_encoder_encode_end(self._h))
        self._source_in_flight = None
        return self.bytes_written()

    def encode_components(self, source, source_component_count: int, stride: int = 0) -> int:
        addr, size, _keep = _as_buffer(source)
        self.lib.check(
            self.lib.charls_jpegls_encoder_encode_components_from_buffer(self._h, addr, size, source_component_count, stride)
        )
        return self.bytes_written()

    def create_abbreviated_format(self):
        self.lib.check(self.lib.charls_jpegls_encoder_create_abbreviated_format(self._h))
        return self.bytes_written()

    def bytes_written(self) -> int:
        n = c_size_t()
        self.lib.check(self.lib.charls_jpegls_encoder_get_bytes_written(self._h, byref(n)))
        return n.value

    def rewind(self):
        self.lib.check(self.lib.charls_jpegls_encoder_rewind(self._h))
        return self


class JpegLSDecoder:
    def __init__(self, lib: CharlsLibrary | None = None):
        self.lib = lib or default_library()
        self._h = self.lib.charls_jpegls_decoder_create()
        if not self._h:
            raise MemoryError("charls_jpegls_decoder_create failed")
        self._source = None
        self._callbacks = []

    def close(self):
        if self._h:
            self.lib.charls_jpegls_decoder_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def source(self, data):
        addr, size, keep = _as_buffer(data)
        self._source = keep
        self.lib.check(self.lib.charls_jpegls_decoder_set_source_buffer(self._h, addr, size))
        return self

    def read_spiff_header(self):
        header = SpiffHeader()
        found = c_int32()
        self.lib.check(self.lib.charls_jpegls_decoder_read_spiff_header(self._h, byref(header), byref(found)))
        return header if found.value else None

    def read_header(self):
        self.lib.check(self.lib.charls_jpegls_decoder_read_header(self._h))
        return self

    def frame_info(self) -> FrameInfo:
        fi = FrameInfo()
        self.lib.check(self.lib.charls_jpegls_decoder_get_frame_info(self._h, byref(fi)))
        return fi

    def near_lossless(self, component_index: int = 0) -> int:
        v = c_int32()
        self.lib.check(self.lib.charls_jpegls_decoder_get_near_lossless(self._h, component_index, byref(v)))
        return v.value

    def interleave_mode(self, component_index: int = 0) -> int:
        v = c_int32()
        self.lib.check(self.lib.charls_jpegls_decoder_get_interleave_mode(self._h, component_index, byref(v)))
        return v.value

    def preset_coding_parameters(self) -> PcParameters:
        pc = PcParameters()
        self.lib.check(self.lib.charls_jpegls_decoder_get_preset_coding_parameters(self._h, 0, byref(pc)))
        return pc

    def color_transformation(self) -> int:
        v = c_int32()
        self.lib.check(self.lib.charls_jpegls_decoder_get_color_transformation(self._h, byref(v)))
        return v.value

    def restart_interval(self) -> int:
        v = c_uint32()
        self.lib.check(self.lib.charlsx_jpegls_decoder_get_restart_interval(self._h, byref(v)))
        return v.value

    def destination_size(self, stride: int = 0) -> int:
        n = c_size_t()
        self.lib.check(self.lib.charls_jpegls_decoder_get_destination_size(self._h, stride, byref(n)))
        return n.value

    def compressed_data_format(self) -> int:
        v = c_int32()
        self.lib.check(self.lib.charls_decoder_get_compressed_data_format(self._h, byref(v)))
        return v.value

    def at_comment(self, fn):
        from .capi import AT_COMMENT_HANDLER

        cb = AT_COMMENT_HANDLER(fn) if fn else AT_COMMENT_HANDLER()
        self._callbacks.append(cb)
        self.lib.check(self.lib.charls_jpegls_decoder_at_comment(self._h, cb, None))
        return self

    def at_application_data(self, fn):
        from .capi import AT_APPLICATION_DATA_HANDLER

        cb = AT_APPLICATION_DATA_HANDLER(fn) if fn else AT_APPLICATION_DATA_HANDLER()
        self._callbacks.append(cb)
        self.lib.check(self.lib.charls_jpegls_decoder_at_application_data(self._h, cb, None))
        return self

    def decode(self, destination: np.ndarray | None = None, stride: int = 0) -> np.ndarray:
        if destination is None:
            destination = np.empty(self.destination_size(stride), dtype=np.uint8)
        addr, size, _keep = _as_buffer(destination)
        self.lib.check(self.lib.charls_jpegls_decoder_decode_to_buffer(self._h, addr, size, stride))
        return destination


    def decode_begin(self, destination: np.ndarray, stride: int = 0):
        addr, size, keep = _as_buffer(destination)
        self._destination_in_flight = keep
        self.lib.check(self.lib.charlsx_jpegls_decoder_decode_to_buffer_begin(self._h, addr, size, stride))
        return self

    def decode_end(self):
        self.lib.check(self.lib.charlsx_jpegls_decoder_decode_end(self._h))
        self._destination_in_flight = None
        return self


# -- convenience helpers ------------------------------------------------------------------------------------------


def _shape_info(image: np.ndarray, interleave_mode: int):
    """(height, width, components) for [H,W] (mono), [H,W,C] (interleaved) or [C,H,W] (planar, ILV none)."""
    if image.ndim == 2:
        return image.shape[0], image.shape[1], 1
    if image.ndim != 3:
        raise ValueError("image must be [H,W], [H,W,C] or (ILV none) [C,H,W]")
    if interleave_mode == INTERLEAVE_NONE:
        return image.shape[1], image.shape[2], image.shape[0]
    return image.shape[0], image.shape[1], image.shape[2]


def encode(
    image: np.ndarray,
    bits_per_sample: int | None = None,
    *,
    near_lossless: int = 0,
    interleave_mode: int = INTERLEAVE_NONE,
    color_transformation: int = TRANSFORM_NONE,
    restart_interval: int | None = None,
    preset: tuple | None = None,
    lib: CharlsLibrary | None = None,
) -> bytes:
    """One-call encode of a numpy image through the C ABI (host buffers in, JPEG-LS bytes out)."""
    lib = lib or default_library()
    if bits_per_sample is None:
        bits_per_sample = 8 if image.dtype == np.uint8 else 16
    h, w, c = _shape_info(image, interleave_mode)
    with JpegLSEncoder(lib) as enc:
        enc.frame_info(w, h, bits_per_sample, c).near_lossless(near_lossless).interleave_mode(interleave_mode)
        enc.color_transformation(color_transformation)
        if preset is not None:
            enc.preset_coding_parameters(*preset)
        if restart_interval is not None:
            enc.restart_interval(restart_interval)
        dst = np.empty(enc.estimated_destination_size(), dtype=np.uint8)
        enc.destination(dst)
        n = enc.encode(image)
        return dst[:n].tobytes()


def decode(stream, *, lib: CharlsLibrary | None = None):
    """One-call decode; returns (pixels, frame_info, interleave_mode). Pixels are shaped like `encode` expects."""
    lib = lib or default_library()
    with JpegLSDecoder(lib) as dec:
        dec.source(stream).read_header()
        fi = dec.frame_info()
        ilv = dec.interleave_mode(0)
        raw = dec.decode()
    dtype = np.uint8 if fi.bits_per_sample <= 8 else np.dtype("<u2")
    px = raw.view(dtype)
    if fi.component_count == 1:
        px = px.reshape(fi.height, fi.width)
    elif ilv == INTERLEAVE_NONE:
        px = px.reshape(fi.component_count, fi.height, fi.width)
    else:
        px = px.reshape(fi.height, fi.width, fi.component_count)
    return px, fi, ilv
