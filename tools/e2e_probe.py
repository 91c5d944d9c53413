"""Where does the end-to-end time go?  Per-call latency of the C-ABI encode / decode calls with pinned host buffers at
several thread counts (cfg2 frames).  Prints one line per thread count."""
import ctypes as C
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from charls_b200 import capi  # noqa: E402
from charls_b200.capi import FrameInfo  # noqa: E402

lib = capi.default_library()
w = h = 4096
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
frames = bench.make_frames(torch, torch.device("cuda"), n, "cfg2", 1234)
frames_host = frames.cpu().pin_memory()
cap = w * h * 2 + 65536
streams_host = torch.empty((n, cap), dtype=torch.uint8).pin_memory()
out_host = torch.empty_like(frames_host).pin_memory()
sizes = [0] * n
t_enc = [0.0] * n
t_dec = [0.0] * n
frame_bytes = w * h


def enc(i):
    t0 = time.perf_counter()
    e = lib.charls_jpegls_encoder_create()
    fi = FrameInfo(w, h, 8, 1)
    lib.check(lib.charls_jpegls_encoder_set_frame_info(e, C.byref(fi)))
    lib.check(lib.charls_jpegls_encoder_set_destination_buffer(e, streams_host[i].data_ptr(), cap))
    lib.check(lib.charls_jpegls_encoder_encode_from_buffer(e, frames_host[i].data_ptr(), frame_bytes, 0))
    written = C.c_size_t()
    lib.check(lib.charls_jpegls_encoder_get_bytes_written(e, C.byref(written)))
    lib.charls_jpegls_encoder_destroy(e)
    sizes[i] = written.value
    t_enc[i] = time.perf_counter() - t0


def dec(i):
    t0 = time.perf_counter()
    d = lib.charls_jpegls_decoder_create()
    lib.check(lib.charls_jpegls_decoder_set_source_buffer(d, streams_host[i].data_ptr(), sizes[i]))
    lib.check(lib.charls_jpegls_decoder_read_header(d))
    lib.check(lib.charls_jpegls_decoder_decode_to_buffer(d, out_host[i].data_ptr(), frame_bytes, 0))
    lib.charls_jpegls_decoder_destroy(d)
    t_dec[i] = time.perf_counter() - t0


def both(i):
    enc(i)
    dec(i)


for threads in ([int(t) for t in sys.argv[2].split(',')] if len(sys.argv) > 2 else (1, 4, 8, 16, 32)):
    for rep in range(3):
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=threads) as pool:
            list(pool.map(both, range(n)))
        dt = time.perf_counter() - t0
    print(f"threads {threads:2d}: {n * w * h / dt / 1e9:6.2f} GPix/s  step {dt * 1e3:7.2f} ms  "
          f"enc call {sum(t_enc) / n * 1e3:6.2f} ms  dec call {sum(t_dec) / n * 1e3:6.2f} ms  ratio {w * h / (sum(sizes) / n):.2f}")
assert torch.equal(out_host, frames_host)

if len(sys.argv) > 3 and sys.argv[3] == "python-only":
    sys.exit(0)

# ---- the same round trips from C++ threads (charls_b200/csrc/driver): no interpreter lock between the ABI calls
from charls_b200 import driver  # noqa: E402

for threads in (1, 4, 8, 16, 32):
    for rep in range(3):
        out_host.zero_()
        dt, sizes_c = driver.run_round_trips(capi.DEFAULT_LIBRARY, frames_host.data_ptr(), frame_bytes, streams_host.data_ptr(), cap,
                                            out_host.data_ptr(), n, threads, width=w, height=h, bits_per_sample=8)
    assert torch.equal(out_host, frames_host)
    print(f"C++ driver, threads {threads:2d}: {n * w * h / dt / 1e9:6.2f} GPix/s  step {dt * 1e3:7.2f} ms")

# ---- host buffers through the host-batch interface: one call per direction, chunks staged by the library
from charls_b200.batch import BatchCodec  # noqa: E402

hb = BatchCodec(w, h, 8)
hb_streams = [streams_host[i] for i in range(n)]
hb_frames = [frames_host[i] for i in range(n)]
hb_out = [out_host[i] for i in range(n)]
for rep in range(3):
    out_host.zero_()
    t0 = time.perf_counter()
    hb_sizes = hb.encode_host(hb_frames, hb_streams)
    t1 = time.perf_counter()
    hb.decode_host(hb_streams, hb_sizes, hb_out)
    t2 = time.perf_counter()
assert torch.equal(out_host, frames_host)
print(f"host-batch interface, 1 thread: encode {n * w * h / (t1 - t0) / 1e9:6.2f} GPix/s, decode {n * w * h / (t2 - t1) / 1e9:6.2f} GPix/s, "
      f"encode then decode {n * w * h / (t2 - t0) / 1e9:6.2f} GPix/s")
# two batches on two threads, one encoding while the other decodes: both PCIe directions busy
half = n // 2
hb2 = BatchCodec(w, h, 8)


def pass_a():
    s = hb.encode_host(hb_frames[:half], hb_streams[:half])
    hb.decode_host(hb_streams[:half], s, hb_out[:half])


def pass_b():
    s = hb2.encode_host(hb_frames[half:], hb_streams[half:])
    hb2.decode_host(hb_streams[half:], s, hb_out[half:])


for rep in range(3):
    out_host.zero_()
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=2) as pool:
        list(pool.map(lambda f: f(), (pass_a, pass_b)))
    dt = time.perf_counter() - t0
assert torch.equal(out_host, frames_host)
print(f"host-batch interface, 2 threads (each: encode half, decode half): {n * w * h / dt / 1e9:6.2f} GPix/s")

# ---- the same round trips with frames and streams resident in HBM (no PCIe): one BatchCodec and CUDA stream per thread
from charls_b200.batch import BatchCodec  # noqa: E402

frames_dev = frames
for threads in (1, 4, 8, 16):
    codecs = [BatchCodec(w, h, 8) for _ in range(threads)]
    cuda_streams = [torch.cuda.Stream() for _ in range(threads)]
    bufs = [torch.empty((1, codecs[0].stream_capacity), dtype=torch.uint8, device="cuda") for _ in range(threads)]
    outs = [torch.empty((1, h, w), dtype=torch.uint8, device="cuda") for _ in range(threads)]

    def device_round_trip(k):
        for i in range(k, n, threads):
            sz = codecs[k].encode(frames_dev[i : i + 1], bufs[k], stream=cuda_streams[k])
            codecs[k].decode(bufs[k], sz, outs[k], stream=cuda_streams[k])
        cuda_streams[k].synchronize()

    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=threads) as pool:
            list(pool.map(device_round_trip, range(threads)))
        dt = time.perf_counter() - t0
    print(f"device-resident, threads {threads:2d}: {n * w * h / dt / 1e9:6.2f} GPix/s  {n / dt:7.0f} frames/s")
