"""Device-resident batch interface (charlsx_batch_*): frames and streams stay in HBM, only headers cross PCIe.

PyTorch is used for what it is good at here -- device memory and streams; the codec itself is the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from ctypes import byref

from .capi import BatchImage, BatchParams, CharlsLibrary, FrameInfo, default_library


class BatchCodec:
    """Encodes / decodes batches of equally shaped frames that live in CUDA tensors."""

    def __init__(self, width, height, bits_per_sample, component_count=1, *, near_lossless=0, interleave_mode=0,
                 color_transformation=0, restart_interval=1, offset_table=False, row_stride=0, lib: CharlsLibrary | None = None):
        """offset_table: write (encode) / look for (decode) the side table of interval offsets in the streams' headers
        (CHARLSX_BATCH_OFFSET_TABLE, include/charls_b200.h).  row_stride: bytes between the lines of a frame (0 = tightly packed);
        device frames whose stride is a multiple of 4 take the tile kernels whatever their width."""
        self.lib = lib or default_library()
        self.params = BatchParams(
            FrameInfo(width, height, bits_per_sample, component_count), near_lossless, interleave_mode, color_transformation,
            restart_interval, row_stride, 1 if offset_table and restart_interval else 0,
        )
        self._h = self.lib.charlsx_batch_create()
        if not self._h:
            raise MemoryError("charlsx_batch_create failed")
        sample_bytes = 1 if bits_per_sample <= 8 else 2
        self.frame_bytes = width * height * component_count * sample_bytes
        # same bound the single-image encoder reports (reference formula + restart-marker overhead)
        intervals = (height + restart_interval - 1) // restart_interval if restart_interval else 0
        self.stream_capacity = self.frame_bytes + self.frame_bytes // 16 + 1024 + 34 + 4 * intervals + 8
        if offset_table and restart_interval:
            self.stream_capacity += 4 * (intervals + 1) + 26 * ((intervals + 1) // 16377 + 1)
        self.stream_capacity = (self.stream_capacity + 255) // 256 * 256

    def close(self):
        if self._h:
            self.lib.charlsx_batch_destroy(self._h)
            self._h = None

    __del__ = close

    _IMAGE_DTYPE = None

    def _images(self, pixels, streams, sizes):
        """The charlsx_batch_image array of a call, filled column by column through a numpy view of the same memory (a Python
        loop over 128 frames costs more than the library needs to issue the whole batch)."""
        import numpy as np

        if BatchCodec._IMAGE_DTYPE is None:
            BatchCodec._IMAGE_DTYPE = np.dtype([("pixels", "<u8"), ("stream", "<u8"), ("stream_capacity", "<u8"), ("stream_size", "<u8"),
                                                ("status", "<i4"), ("reserved", "<i4")])
            assert BatchCodec._IMAGE_DTYPE.itemsize == C.sizeof(BatchImage)
        n = pixels.shape[0]
        images = (BatchImage * n)()
        view = np.frombuffer(images, dtype=BatchCodec._IMAGE_DTYPE)
        index = np.arange(n, dtype=np.uint64)
        view["pixels"] = pixels.data_ptr() + index * np.uint64(pixels.stride(0) * pixels.element_size())
        stream_stride = streams.stride(0) * streams.element_size()
        view["stream"] = streams.data_ptr() + index * np.uint64(stream_stride)
        view["stream_capacity"] = stream_stride if sizes is None else np.asarray(sizes, dtype=np.uint64)
        self._last_view = view
        return images

    def _stream_sizes(self, images):
        return self._last_view["stream_size"].tolist()

    @staticmethod
    def _stream_handle(stream):
        if stream is None:
            import torch

            stream = torch.cuda.current_stream()
        return C.c_void_p(stream.cuda_stream)

    def encode(self, pixels, streams, stream=None):
        """pixels: CUDA tensor [N, ...frame]; streams: CUDA uint8 tensor [N, capacity].  Returns list of stream sizes."""
        images = self._images(pixels, streams, None)
        errc = self.lib.charlsx_batch_encode(self._h, byref(self.params), images, len(images), self._stream_handle(stream))
        self.lib.check(errc)
        return self._stream_sizes(images)

    def decode(self, streams, sizes, pixels, stream=None):
        """streams: CUDA uint8 tensor [N, capacity] holding complete JPEG-LS streams of `sizes` bytes; pixels: output."""
        images = self._images(pixels, streams, sizes)
        errc = self.lib.charlsx_batch_decode(self._h, byref(self.params), images, len(images), self._stream_handle(stream))
        self.lib.check(errc)
        return self._stream_sizes(images)

    def decode_new(self, streams, sizes, stream=None):
        """decode() into a new CUDA tensor [N, H, W] or [N, H, W, C] (uint8, or int16 carrying the uint16 bit pattern)."""
        import torch

        fi = self.params.frame_info
        shape = (streams.shape[0], fi.height, fi.width) + ((fi.component_count,) if fi.component_count > 1 else ())
        out = torch.empty(shape, dtype=torch.uint8 if fi.bits_per_sample <= 8 else torch.int16, device=streams.device)
        self.decode(streams, sizes, out, stream)
        return out

    # ---- frames and streams in host memory (numpy arrays or pinned torch tensors); see charlsx_batch_encode_host
    @staticmethod
    def _host_images(pixels, streams, sizes):
        n = len(pixels)
        images = (BatchImage * n)()
        for i in range(n):
            images[i].pixels = pixels[i].ctypes.data if hasattr(pixels[i], "ctypes") else pixels[i].data_ptr()
            images[i].stream = streams[i].ctypes.data if hasattr(streams[i], "ctypes") else streams[i].data_ptr()
            capacity = streams[i].nbytes if hasattr(streams[i], "nbytes") else streams[i].numel() * streams[i].element_size()
            images[i].stream_capacity = capacity if sizes is None else int(sizes[i])
        return images

    def encode_host(self, pixels, streams):
        """pixels[i], streams[i]: host arrays (frame i, its stream buffer).  Returns the list of stream sizes."""
        images = self._host_images(pixels, streams, None)
        self.lib.check(self.lib.charlsx_batch_encode_host(self._h, byref(self.params), images, len(images)))
        return [images[i].stream_size for i in range(len(images))]

    def decode_host(self, streams, sizes, pixels):
        """streams[i]: host array with a complete JPEG-LS stream of sizes[i] bytes; pixels[i]: host output array."""
        images = self._host_images(pixels, streams, sizes)
        self.lib.check(self.lib.charlsx_batch_decode_host(self._h, byref(self.params), images, len(images)))
        return [images[i].stream_size for i in range(len(images))]

    def last_coder_kernel_ms(self) -> float:
        ms = C.c_float()
        self.lib.check(self.lib.charlsx_batch_get_last_coder_kernel_ms(self._h, byref(ms)))
        return ms.value

    def last_kernel_launches(self) -> int:
        n = C.c_uint32()
        self.lib.check(self.lib.charlsx_batch_get_last_kernel_launches(self._h, byref(n)))
        return n.value
