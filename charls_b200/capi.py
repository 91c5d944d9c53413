"""ctypes binding of the CharLS C ABI (the 48 ``charls_*`` symbols) plus the ``charlsx_*`` B200 extensions.

The binding is ABI-generic on purpose: ``CharlsLibrary(path)`` works for any shared library that exports the
reference's C interface (reference: include/charls/charls_jpegls_encoder.h:24-316,
include/charls/charls_jpegls_decoder.h:24-293, include/charls/public_types.h:934-1034).  The product loads
``charls_b200/lib/libcharls.so.3`` (CUDA engine); the test-suite additionally points the same class at the
unmodified reference build to compare the two libraries call-for-call.
"""
from __future__ import annotations

import ctypes as C
import os
from ctypes import POINTER, byref, c_char_p, c_int32, c_size_t, c_uint32, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# CHARLS_B200_LIBRARY points the package at another build of the same library (tools/ab_build.py: kernel experiments)
DEFAULT_LIBRARY = os.environ.get("CHARLS_B200_LIBRARY") or os.path.join(_HERE, "lib", "libcharls.so.3")


class FrameInfo(C.Structure):
    """charls_frame_info (reference public_types.h:988-1001), 16 bytes."""

    _fields_ = [("width", c_uint32), ("height", c_uint32), ("bits_per_sample", c_int32), ("component_count", c_int32)]


class PcParameters(C.Structure):
    """charls_jpegls_pc_parameters (reference public_types.h:1008-1020), 20 bytes."""

    _fields_ = [
        ("maximum_sample_value", c_int32),
        ("threshold1", c_int32),
        ("threshold2", c_int32),
        ("threshold3", c_int32),
        ("reset_value", c_int32),
    ]


class SpiffHeader(C.Structure):
    """charls_spiff_header (reference public_types.h:934-950), 40 bytes."""

    _fields_ = [
        ("profile_id", c_int32),
        ("component_count", c_int32),
        ("height", c_uint32),
        ("width", c_uint32),
        ("color_space", c_int32),
        ("bits_per_sample", c_int32),
        ("compression_type", c_int32),
        ("resolution_units", c_int32),
        ("vertical_resolution", c_uint32),
        ("horizontal_resolution", c_uint32),
    ]


class MappingTableInfo(C.Structure):
    """charls_mapping_table_info (reference public_types.h:1027-1034), 12 bytes."""

    _fields_ = [("table_id", c_int32), ("entry_size", c_int32), ("data_size", c_uint32)]


AT_COMMENT_HANDLER = C.CFUNCTYPE(c_int32, c_void_p, c_size_t, c_void_p)
AT_APPLICATION_DATA_HANDLER = C.CFUNCTYPE(c_int32, c_int32, c_void_p, c_size_t, c_void_p)

# name -> (restype, argtypes); the complete reference ABI.
_E = c_void_p  # charls_jpegls_encoder*
_D = c_void_p  # charls_jpegls_decoder*
_ERR = c_int32
ABI_SYMBOLS = {
    # encoder (reference include/charls/charls_jpegls_encoder.h)
    "charls_jpegls_encoder_create": (_E, []),
    "charls_jpegls_encoder_destroy": (None, [_E]),
    "charls_jpegls_encoder_set_frame_info": (_ERR, [_E, POINTER(FrameInfo)]),
    "charls_jpegls_encoder_set_near_lossless": (_ERR, [_E, c_int32]),
    "charls_jpegls_encoder_set_encoding_options": (_ERR, [_E, c_uint32]),
    "charls_jpegls_encoder_set_interleave_mode": (_ERR, [_E, c_int32]),
    "charls_jpegls_encoder_set_preset_coding_parameters": (_ERR, [_E, POINTER(PcParameters)]),
    "charls_jpegls_encoder_set_color_transformation": (_ERR, [_E, c_int32]),
    "charls_jpegls_encoder_set_mapping_table_id": (_ERR, [_E, c_int32, c_int32]),
    "charls_jpegls_encoder_get_estimated_destination_size": (_ERR, [_E, POINTER(c_size_t)]),
    "charls_jpegls_encoder_set_destination_buffer": (_ERR, [_E, c_void_p, c_size_t]),
    "charls_jpegls_encoder_write_standard_spiff_header": (_ERR, [_E, c_int32, c_int32, c_uint32, c_uint32]),
    "charls_jpegls_encoder_write_spiff_header": (_ERR, [_E, POINTER(SpiffHeader)]),
    "charls_jpegls_encoder_write_spiff_entry": (_ERR, [_E, c_uint32, c_void_p, c_size_t]),
    "charls_jpegls_encoder_write_spiff_end_of_directory_entry": (_ERR, [_E]),
    "charls_jpegls_encoder_write_comment": (_ERR, [_E, c_void_p, c_size_t]),
    "charls_jpegls_encoder_write_application_data": (_ERR, [_E, c_int32, c_void_p, c_size_t]),
    "charls_jpegls_encoder_write_mapping_table": (_ERR, [_E, c_int32, c_int32, c_void_p, c_size_t]),
    "charls_jpegls_encoder_encode_from_buffer": (_ERR, [_E, c_void_p, c_size_t, c_uint32]),
    "charls_jpegls_encoder_encode_components_from_buffer": (_ERR, [_E, c_void_p, c_size_t, c_int32, c_uint32]),
    "charls_jpegls_encoder_create_abbreviated_format": (_ERR, [_E]),
    "charls_jpegls_encoder_get_bytes_written": (_ERR, [_E, POINTER(c_size_t)]),
    "charls_jpegls_encoder_rewind": (_ERR, [_E]),
    # decoder (reference include/charls/charls_jpegls_decoder.h)
    "charls_jpegls_decoder_create": (_D, []),
    "charls_jpegls_decoder_destroy": (None, [_D]),
    "charls_jpegls_decoder_set_source_buffer": (_ERR, [_D, c_void_p, c_size_t]),
    "charls_jpegls_decoder_read_spiff_header": (_ERR, [_D, POINTER(SpiffHeader), POINTER(c_int32)]),
    "charls_jpegls_decoder_read_header": (_ERR, [_D]),
    "charls_jpegls_decoder_get_frame_info": (_ERR, [_D, POINTER(FrameInfo)]),
    "charls_jpegls_decoder_get_near_lossless": (_ERR, [_D, c_int32, POINTER(c_int32)]),
    "charls_jpegls_decoder_get_interleave_mode": (_ERR, [_D, c_int32, POINTER(c_int32)]),
    "charls_jpegls_decoder_get_preset_coding_parameters": (_ERR, [_D, c_int32, POINTER(PcParameters)]),
    "charls_jpegls_decoder_get_color_transformation": (_ERR, [_D, POINTER(c_int32)]),
    "charls_jpegls_decoder_get_destination_size": (_ERR, [_D, c_uint32, POINTER(c_size_t)]),
    "charls_jpegls_decoder_decode_to_buffer": (_ERR, [_D, c_void_p, c_size_t, c_uint32]),
    "charls_jpegls_decoder_at_comment": (_ERR, [_D, AT_COMMENT_HANDLER, c_void_p]),
    "charls_jpegls_decoder_at_application_data": (_ERR, [_D, AT_APPLICATION_DATA_HANDLER, c_void_p]),
    "charls_decoder_get_compressed_data_format": (_ERR, [_D, POINTER(c_int32)]),
    "charls_decoder_get_mapping_table_id": (_ERR, [_D, c_int32, POINTER(c_int32)]),
    "charls_decoder_find_mapping_table_index": (_ERR, [_D, c_int32, POINTER(c_int32)]),
    "charls_decoder_get_mapping_table_count": (_ERR, [_D, POINTER(c_int32)]),
    "charls_decoder_get_mapping_table_info": (_ERR, [_D, c_int32, POINTER(MappingTableInfo)]),
    "charls_decoder_get_mapping_table_data": (_ERR, [_D, c_int32, c_void_p, c_size_t]),
    # misc (reference jpegls_error.h:12, jpegls_error.hpp:10, version.h:27-38, validate_spiff_header.h:23-24)
    "charls_get_error_message": (c_char_p, [_ERR]),
    "charls_get_jpegls_category": (c_void_p, []),
    "charls_get_version_string": (c_char_p, []),
    "charls_get_version_number": (None, [POINTER(c_int32), POINTER(c_int32), POINTER(c_int32)]),
    "charls_validate_spiff_header": (_ERR, [POINTER(SpiffHeader), POINTER(FrameInfo)]),
}


class BatchImage(C.Structure):
    """charlsx_batch_image (include/charls_b200.h): one frame of a device-resident batch."""

    _fields_ = [
        ("pixels", c_void_p),  # device pointer: raw samples (encode: in, decode: out)
        ("stream", c_void_p),  # device pointer: JPEG-LS byte stream (encode: out, decode: in)
        ("stream_capacity", c_size_t),  # encode: capacity of `stream`; decode: size of the stream in bytes
        ("stream_size", c_size_t),  # encode: bytes written (out)
        ("status", c_int32),  # charls_jpegls_errc of this frame (out)
        ("reserved", c_int32),
    ]


class BatchParams(C.Structure):
    """charlsx_batch_params (include/charls_b200.h): geometry/coding parameters shared by all frames of a batch."""

    _fields_ = [
        ("frame_info", FrameInfo),
        ("near_lossless", c_int32),
        ("interleave_mode", c_int32),
        ("color_transformation", c_int32),
        ("restart_interval", c_uint32),
        ("stride", c_uint32),
        ("flags", c_uint32),
    ]


# B200 extensions (not in the reference; include/charls_b200.h documents each one)
EXT_SYMBOLS = {
    "charlsx_get_device_count": (_ERR, [POINTER(c_int32)]),
    "charlsx_set_device": (_ERR, [c_int32]),
    "charlsx_jpegls_encoder_set_restart_interval": (_ERR, [_E, c_uint32]),
    "charlsx_jpegls_decoder_get_restart_interval": (_ERR, [_D, POINTER(c_uint32)]),
    "charlsx_jpegls_encoder_set_offset_table": (_ERR, [_E, c_int32]),
    "charlsx_jpegls_encoder_encode_from_buffer_begin": (_ERR, [_E, c_void_p, c_size_t, c_uint32]),
    "charlsx_jpegls_encoder_encode_end": (_ERR, [_E]),
    "charlsx_jpegls_decoder_decode_to_buffer_begin": (_ERR, [_D, c_void_p, c_size_t, c_uint32]),
    "charlsx_jpegls_decoder_decode_end": (_ERR, [_D]),
    "charlsx_batch_create": (c_void_p, []),
    "charlsx_batch_destroy": (None, [c_void_p]),
    "charlsx_batch_encode": (_ERR, [c_void_p, POINTER(BatchParams), POINTER(BatchImage), c_size_t, c_void_p]),
    "charlsx_batch_decode": (_ERR, [c_void_p, POINTER(BatchParams), POINTER(BatchImage), c_size_t, c_void_p]),
    "charlsx_batch_encode_host": (_ERR, [c_void_p, POINTER(BatchParams), POINTER(BatchImage), c_size_t]),
    "charlsx_batch_decode_host": (_ERR, [c_void_p, POINTER(BatchParams), POINTER(BatchImage), c_size_t]),
    "charlsx_batch_get_last_kernel_launches": (_ERR, [c_void_p, POINTER(c_uint32)]),
    "charlsx_batch_get_last_coder_kernel_ms": (_ERR, [c_void_p, POINTER(C.c_float)]),
    "charlsx_get_kernel_launch_count": (_ERR, [POINTER(C.c_uint64)]),
}


class CharlsError(RuntimeError):
    """Raised for a non-zero charls_jpegls_errc; mirrors charls::jpegls_error (reference jpegls_error.hpp:34-60)."""

    def __init__(self, errc: int, message: str):
        super().__init__(f"jpegls_errc {errc}: {message}")
        self.errc = errc


class CharlsLibrary:
    """A loaded libcharls-compatible shared library with typed entry points."""

    def __init__(self, path: str | None = None, *, extensions: bool | None = None):
        self.path = path or DEFAULT_LIBRARY
        if not os.path.exists(self.path):
            raise FileNotFoundError(
                f"{self.path} not found - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)"
            )
        self.dll = C.CDLL(self.path)
        for name, (restype, argtypes) in ABI_SYMBOLS.items():
            fn = getattr(self.dll, name)
            fn.restype = restype
            fn.argtypes = argtypes
            setattr(self, name, fn)
        self.has_extensions = hasattr(self.dll, "charlsx_batch_encode") if extensions is None else extensions
        if self.has_extensions:
            for name, (restype, argtypes) in EXT_SYMBOLS.items():
                fn = getattr(self.dll, name)
                fn.restype = restype
                fn.argtypes = argtypes
                setattr(self, name, fn)

    def check(self, errc: int) -> None:
        if errc != 0:
            msg = self.charls_get_error_message(errc)
            raise CharlsError(errc, msg.decode("utf-8", "replace") if msg else "")

    def version_string(self) -> str:
        return self.charls_get_version_string().decode()


_default = None


def default_library() -> CharlsLibrary:
    """The in-tree CUDA build of the ABI.  Fails loudly if it has not been built."""
    global _default
    if _default is None:
        _default = CharlsLibrary(DEFAULT_LIBRARY)
    return _default


__all__ = [
    "ABI_SYMBOLS",
    "EXT_SYMBOLS",
    "AT_APPLICATION_DATA_HANDLER",
    "AT_COMMENT_HANDLER",
    "BatchImage",
    "BatchParams",
    "CharlsError",
    "CharlsLibrary",
    "DEFAULT_LIBRARY",
    "FrameInfo",
    "MappingTableInfo",
    "PcParameters",
    "SpiffHeader",
    "byref",
    "default_library",
]
