#!/usr/bin/env python
"""bench.py -- MPixels/s of the JPEG-LS scan engine (encode + decode), HBM roofline fraction, CPU baseline.

    python bench.py --gpus N --steps K --warmup W            # this repository's CUDA engine
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU codec on the box's host cores

One *step* = encode + decode of `--frames` synthetic frames per GPU with restart interval 1 (every line an independent
work item).  Frames and streams are resident in HBM when the timed region starts (`value`); `e2e` measures the same
round trip through the reference-facing C ABI (charls_jpegls_encoder_encode_from_buffer /
charls_jpegls_decoder_decode_to_buffer) with pinned HOST buffers, copies included.
Multi GPU (torchrun): frames are sharded across ranks, no data-path collective; NCCL only all-gathers the stream sizes.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time
from collections import deque
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (width, height, bits, components, near, interleave, transform)
    "cfg2": (4096, 4096, 8, 1, 0, 0, 0),   # BASELINE.json configs[1]: 4096x4096 8-bit grayscale lossless
    "cfg3": (4096, 4096, 12, 1, 2, 0, 0),  # configs[2]: 12-bit NEAR=2
    "cfg4": (2048, 2048, 16, 3, 0, 2, 1),  # configs[3]: 16-bit RGB, ILV sample, HP1
    # not in BASELINE.json: line interleave (kernels without shared-memory tiles, k_encode_fast / k_decode_fast) and rows that
    # do not start on word boundaries (copied to an aligned pitch around the tile kernels)
    "rgb8line": (2048, 2048, 8, 3, 0, 1, 0),  # line interleave: a lane gathers its component from RGBRGB... rows
    "odd8": (4095, 4096, 8, 1, 0, 0, 0),      # rows of 4095 bytes, tightly packed (--row-pitch 4096: no copy needed)
}
METRIC = "MPixels/s encode+decode"
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libcharls_ref.so")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=128, help="frames per GPU per step (128 x 8 GPUs = BASELINE configs[4])")
    ap.add_argument("--e2e-frames", type=int, default=32)
    ap.add_argument("--e2e-threads", type=int, default=8)
    ap.add_argument("--e2e-one-part-threads", type=int, default=16, help="host threads of the headline e2e run (one image per thread)")
    ap.add_argument("--e2e-inflight", type=int, default=4,
                    help="codec objects every e2e host thread keeps in flight through the two-part calls (charlsx_*_begin / _end); "
                         "1 = the one-part reference calls only")
    ap.add_argument("--e2e-passes", type=int, default=4,
                    help="every pinned frame buffer makes this many round trips per e2e step (32 x 4 = 128 frames per step)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-phases", action="store_true", help="e2e: encode all frames, then decode all (default: per frame)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--also", default="auto", choices=["auto", "none"] + sorted(WORKLOADS),
                    help="second workload reported under \"also\" (auto: cfg4 = 16-bit RGB beside cfg2, the metric names both)")
    ap.add_argument("--restart-interval", type=int, default=1,
                    help="lines per restart interval; 1 = every line an independent work item (the design point), 0 = none: the "
                         "streams the reference itself writes, one CUDA thread per scan (the general path)")
    ap.add_argument("--no-offset-table", action="store_true",
                    help="batch streams without the side table of interval offsets (APP11 \"JLS-OFFT\", include/charls_b200.h): the "
                         "decoder then searches every stream for its restart markers (three more kernels)")
    ap.add_argument("--row-pitch", type=int, default=0,
                    help="bytes between the lines of the device frames (0 = tightly packed); e.g. --workload odd8 --row-pitch 4096")
    ap.add_argument("--content", default="smooth", choices=["smooth", "noise", "flat"],
                    help="smooth = S_smooth (the metric's input); noise (uniform, incompressible) and flat (all zero, pure run "
                         "mode) bracket it (SURVEY.md 8d)")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------------
# synthetic frames: S_smooth of SURVEY.md 8(d) -- smooth base + 1% gaussian noise, generated on the device
# ---------------------------------------------------------------------------------------------------------------------
def make_frames(torch, device, count, workload, first_seed, content="smooth"):
    w, h, bits, cc, _, ilv, _ = WORKLOADS[workload]
    mx = (1 << bits) - 1
    if content != "smooth":
        dtype = torch.uint8 if bits <= 8 else torch.int16
        shape = (count, h, w) if cc == 1 else (count, h, w, cc)
        if content == "flat":
            return torch.zeros(shape, device=device, dtype=dtype)
        gen = torch.Generator(device=device)
        gen.manual_seed(first_seed)
        v = torch.randint(0, mx + 1, shape, device=device, generator=gen, dtype=torch.int32)
        if bits > 8:
            v = torch.where(v > 32767, v - 65536, v)
        return v.to(dtype)
    x = torch.arange(w, device=device, dtype=torch.float32)[None, :]
    y = torch.arange(h, device=device, dtype=torch.float32)[:, None]
    base = 0.8 * mx * (0.5 + 0.25 * torch.sin(x / 97.0) + 0.25 * torch.cos(y / 131.0))
    dtype = torch.uint8 if bits <= 8 else torch.int16  # int16 carries the uint16 bit pattern
    shape = (count, h, w) if cc == 1 else (count, h, w, cc)
    frames = torch.empty(shape, device=device, dtype=dtype)
    gen = torch.Generator(device=device)
    for i in range(count):
        gen.manual_seed(first_seed + i)
        for c in range(cc):
            noise = torch.randn((h, w), device=device, generator=gen) * (0.01 * mx)
            v = torch.clamp(base * (1 - 0.1 * c) + noise, 0, mx).to(torch.int32)
            if bits > 8:
                v = torch.where(v > 32767, v - 65536, v)
            if cc == 1:
                frames[i] = v.to(dtype)
            else:
                frames[i, :, :, c] = v.to(dtype)
    return frames


# ---------------------------------------------------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.proc.wait()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the UNMODIFIED reference (oracle/_ref) on the host cores, one frame per thread
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_step(ref, frames_np, workload, threads, buffers=None):
    """Encodes and decodes frames_np[i] on `threads` host threads with the reference library (objects constructed inside
    the loop, buffers pre-allocated outside, like the reference's own cli/benchmark.cpp:44-89). Returns seconds."""
    from charls_b200.capi import FrameInfo

    w, h, bits, cc, near, ilv, xf = WORKLOADS[workload]
    n = len(frames_np)
    if buffers is None:
        buffers = {}
    if "dst" not in buffers:
        cap = frames_np[0].nbytes + frames_np[0].nbytes // 16 + 2048
        buffers["dst"] = [np.zeros(cap, np.uint8) for _ in range(n)]
        buffers["out"] = [np.zeros(frames_np[0].nbytes, np.uint8) for _ in range(n)]

    def work(i):
        src, dst, out = frames_np[i], buffers["dst"][i], buffers["out"][i]
        e = ref.charls_jpegls_encoder_create()
        fi = FrameInfo(w, h, bits, cc)
        ref.check(ref.charls_jpegls_encoder_set_frame_info(e, C.byref(fi)))
        ref.check(ref.charls_jpegls_encoder_set_near_lossless(e, near))
        ref.check(ref.charls_jpegls_encoder_set_interleave_mode(e, ilv))
        ref.check(ref.charls_jpegls_encoder_set_color_transformation(e, xf))
        ref.check(ref.charls_jpegls_encoder_set_destination_buffer(e, dst.ctypes.data, dst.nbytes))
        ref.check(ref.charls_jpegls_encoder_encode_from_buffer(e, src.ctypes.data, src.nbytes, 0))
        written = C.c_size_t()
        ref.check(ref.charls_jpegls_encoder_get_bytes_written(e, C.byref(written)))
        ref.charls_jpegls_encoder_destroy(e)
        d = ref.charls_jpegls_decoder_create()
        ref.check(ref.charls_jpegls_decoder_set_source_buffer(d, dst.ctypes.data, written.value))
        ref.check(ref.charls_jpegls_decoder_read_header(d))
        ref.check(ref.charls_jpegls_decoder_decode_to_buffer(d, out.ctypes.data, out.nbytes, 0))
        ref.charls_jpegls_decoder_destroy(d)

    workers = [threading.Thread(target=work, args=(i,)) for i in range(n)]
    t0 = time.perf_counter()
    for t in workers:
        t.start()
    for t in workers:
        t.join()
    return time.perf_counter() - t0


def effective_cpus():
    """CPUs this process may really use: min(os.cpu_count(), scheduler affinity, cgroup CPU quota)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period))))
    except (OSError, ValueError):
        pass
    return n


KERNEL_SOURCES = ("jls_kernels.cu", "jls_fast.cuh", "jls_codec.cuh", "jls_tile.cuh", "jls_interval.cuh", "jls_common.h", "jls_params.hpp")


def kernel_sources_sha():
    """Identifies the kernel code a profile was taken with (tools/ncu_traffic.py stores it beside the measured traffic)."""
    import hashlib

    digest = hashlib.sha256()
    for name in KERNEL_SOURCES:
        with open(os.path.join(ROOT, "charls_b200", "csrc", name), "rb") as f:
            digest.update(f.read())
    return digest.hexdigest()[:16]


def host_frames(workload, count, seed=1234):
    """numpy S_smooth frames for the CPU arm (cheap generator shared by `count` frames with different noise)."""
    w, h, bits, cc, _, _, _ = WORKLOADS[workload]
    mx = (1 << bits) - 1
    x = np.arange(w, dtype=np.float32)[None, :]
    y = np.arange(h, dtype=np.float32)[:, None]
    base = 0.8 * mx * (0.5 + 0.25 * np.sin(x / 97.0) + 0.25 * np.cos(y / 131.0))
    dtype = np.uint8 if bits <= 8 else np.dtype("<u2")
    out = []
    for i in range(count):
        rng = np.random.default_rng(seed + i)
        comps = [np.clip(base * (1 - 0.1 * c) + rng.standard_normal((h, w), dtype=np.float32) * (0.01 * mx), 0, mx).astype(dtype)
                 for c in range(cc)]
        out.append(comps[0] if cc == 1 else np.stack(comps, axis=-1))
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w, h, bits, cc, near, ilv, xf = WORKLOADS[args.workload]
    if not os.path.exists(REF_LIB):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libcharls_ref.so was not built (run __graft_entry__.build() where /root/reference exists)"}))
        return
    from charls_b200.capi import CharlsLibrary

    ref = CharlsLibrary(REF_LIB, extensions=False)
    threads = effective_cpus()
    frames = host_frames(args.workload, min(threads, 8))
    frames = [frames[i % len(frames)] for i in range(threads)]  # one frame per thread per step
    buffers = {}
    for _ in range(args.warmup):
        cpu_reference_step(ref, frames, args.workload, threads, buffers)
    t = 0.0
    for _ in range(args.steps):
        t += cpu_reference_step(ref, frames, args.workload, threads, buffers)
    pixels = threads * w * h * args.steps
    value = pixels / t / 1e6
    sample = f"{threads} host threads x 1 frame of {args.workload} per step (reference encode, no restart markers, + decode of its own stream)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "MPixels/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8" if bits <= 8 else "u16", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w}x{h} {bits}-bit x{cc} NEAR={near} ILV={ilv} HP{xf}, no restart markers (the reference "
                               f"cannot write them), {threads} frames per step = one per host thread",
                   "frames_per_step": threads, "restart_interval": 0},
        "cpu_baseline": {"value": value, "unit": "MPixels/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "MPixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_name(args):
    w, h, bits, cc, near, ilv, xf = WORKLOADS[args.workload]
    content = "" if args.content == "smooth" else f", content={args.content}"
    return (f"{args.workload}: {w}x{h} {bits}-bit x{cc} NEAR={near} ILV={ilv} HP{xf} restart-interval={args.restart_interval}, "
            f"{args.frames} frames per GPU per step{content}")


# ---------------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------------
def e2e_round_trip(lib, frames_host, streams_host, out_host, workload, threads, pipelined=True, passes=1, inflight=1):
    """Through the C ABI with pinned host buffers: every frame buffer is encoded and its stream decoded `passes` times (a
    worker owns buffers t, t + threads, ... so no buffer is ever in two calls at once). Returns (s, sizes)."""
    from charls_b200.capi import FrameInfo

    w, h, bits, cc, near, ilv, xf = WORKLOADS[workload]
    n = frames_host.shape[0]
    sizes = [0] * n
    frame_bytes = frames_host[0].numel() * frames_host.element_size()
    cap = streams_host.shape[1]

    def enc(i):
        e = lib.charls_jpegls_encoder_create()
        fi = FrameInfo(w, h, bits, cc)
        lib.check(lib.charls_jpegls_encoder_set_frame_info(e, C.byref(fi)))
        lib.check(lib.charls_jpegls_encoder_set_near_lossless(e, near))
        lib.check(lib.charls_jpegls_encoder_set_interleave_mode(e, ilv))
        lib.check(lib.charls_jpegls_encoder_set_color_transformation(e, xf))
        lib.check(lib.charls_jpegls_encoder_set_destination_buffer(e, streams_host[i].data_ptr(), cap))
        lib.check(lib.charls_jpegls_encoder_encode_from_buffer(e, frames_host[i].data_ptr(), frame_bytes, 0))
        written = C.c_size_t()
        lib.check(lib.charls_jpegls_encoder_get_bytes_written(e, C.byref(written)))
        lib.charls_jpegls_encoder_destroy(e)
        sizes[i] = written.value

    def dec(i):
        d = lib.charls_jpegls_decoder_create()
        lib.check(lib.charls_jpegls_decoder_set_source_buffer(d, streams_host[i].data_ptr(), sizes[i]))
        lib.check(lib.charls_jpegls_decoder_read_header(d))
        lib.check(lib.charls_jpegls_decoder_decode_to_buffer(d, out_host[i].data_ptr(), frame_bytes, 0))
        lib.charls_jpegls_decoder_destroy(d)

    def both(i):
        enc(i)
        dec(i)

    def begin_enc(i):
        e = lib.charls_jpegls_encoder_create()
        fi = FrameInfo(w, h, bits, cc)
        lib.check(lib.charls_jpegls_encoder_set_frame_info(e, C.byref(fi)))
        lib.check(lib.charls_jpegls_encoder_set_near_lossless(e, near))
        lib.check(lib.charls_jpegls_encoder_set_interleave_mode(e, ilv))
        lib.check(lib.charls_jpegls_encoder_set_color_transformation(e, xf))
        lib.check(lib.charls_jpegls_encoder_set_destination_buffer(e, streams_host[i].data_ptr(), cap))
        lib.check(lib.charlsx_jpegls_encoder_encode_from_buffer_begin(e, frames_host[i].data_ptr(), frame_bytes, 0))
        return e

    def end_enc(i, e):
        lib.check(lib.charlsx_jpegls_encoder_encode_end(e))
        written = C.c_size_t()
        lib.check(lib.charls_jpegls_encoder_get_bytes_written(e, C.byref(written)))
        lib.charls_jpegls_encoder_destroy(e)
        sizes[i] = written.value

    def begin_dec(i):
        d = lib.charls_jpegls_decoder_create()
        lib.check(lib.charls_jpegls_decoder_set_source_buffer(d, streams_host[i].data_ptr(), sizes[i]))
        lib.check(lib.charls_jpegls_decoder_read_header(d))
        lib.check(lib.charlsx_jpegls_decoder_decode_to_buffer_begin(d, out_host[i].data_ptr(), frame_bytes, 0))
        return d

    def end_dec(d):
        lib.check(lib.charlsx_jpegls_decoder_decode_end(d))
        lib.charls_jpegls_decoder_destroy(d)

    def worker(t):
        # every worker encodes a frame and decodes it right away: both PCIe directions carry raw and compressed bytes
        # all the time instead of raw going up in one phase and coming down in the next
        if inflight <= 1:
            for _ in range(passes):
                for i in range(t, n, threads):
                    both(i)
            return
        # two-part calls: up to `inflight` codec objects of this thread are on the device at any time.  Completion is FIFO
        # and a frame's buffers belong to one worker, so a buffer is never written while an earlier operation reads it.
        pending = deque()

        def complete_oldest():
            kind, i, handle = pending.popleft()
            if kind == 0:
                end_enc(i, handle)
                pending.append((1, i, begin_dec(i)))
            else:
                end_dec(handle)

        for _ in range(passes):
            for i in range(t, n, threads):
                while len(pending) >= inflight:
                    complete_oldest()
                pending.append((0, i, begin_enc(i)))
        while pending:
            complete_oldest()

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as pool:
        if pipelined:
            list(pool.map(worker, range(threads)))
        else:
            for _ in range(passes):
                list(pool.map(enc, range(n)))
                list(pool.map(dec, range(n)))
    return time.perf_counter() - t0, sizes


def secondary_workload(torch, dist, lib, device, rank, world, workload, F, steps, peak, offset_table=True):
    """The metric names two inputs (8-bit mono and 16-bit RGB): the other one, measured like `value` (device-resident frames,
    CUDA events on the coder stream, max over ranks) with fewer frames and steps, reported under "also"."""
    from charls_b200.batch import BatchCodec

    w, h, bits, cc, near, ilv, xf = WORKLOADS[workload]
    frames = make_frames(torch, device, F, workload, first_seed=1234 + rank * F)
    codec = BatchCodec(w, h, bits, cc, near_lossless=near, interleave_mode=ilv, color_transformation=xf, restart_interval=1,
                       offset_table=offset_table, lib=lib)
    streams = torch.empty((F, codec.stream_capacity), device=device, dtype=torch.uint8)
    decoded = torch.empty_like(frames)
    raw_bytes = frames[0].numel() * frames.element_size()
    for _ in range(3):
        sizes = codec.encode(frames, streams)
        codec.decode(streams, sizes, decoded)
    torch.cuda.synchronize()
    assert near != 0 or torch.equal(decoded, frames), "round trip mismatch"
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    enc_ms, dec_ms = [], []
    start.record()
    for _ in range(steps):
        sizes = codec.encode(frames, streams)
        enc_ms.append(codec.last_coder_kernel_ms())
        codec.decode(streams, sizes, decoded)
        dec_ms.append(codec.last_coder_kernel_ms())
    stop.record()
    torch.cuda.synchronize()
    t = torch.tensor([start.elapsed_time(stop) / steps, float(np.mean(enc_ms)), float(np.mean(dec_ms))], device=device, dtype=torch.float64)
    comp = torch.tensor([float(sum(sizes))], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(comp, op=dist.ReduceOp.SUM)
    del frames, streams, decoded
    torch.cuda.empty_cache()
    ms, t_enc, t_dec = (float(v) for v in t.tolist())
    comp_per_frame = float(comp.item()) / (world * F)
    pixels = world * F * w * h
    algorithmic = F * (raw_bytes + comp_per_frame)
    return {
        "workload": f"{workload}: {w}x{h} {bits}-bit x{cc} NEAR={near} ILV={ilv} HP{xf} restart-interval=1, {F} frames per GPU per step",
        "value": pixels / (ms * 1e-3) / 1e6, "unit": "MPixels/s", "ms_per_step": ms, "steps": steps, "dtype": "u8" if bits <= 8 else "u16",
        "encode_mpix_s": pixels / (t_enc * 1e-3) / 1e6, "decode_mpix_s": pixels / (t_dec * 1e-3) / 1e6,
        "ratio": raw_bytes / comp_per_frame,
        "roofline_frac": {"encode": algorithmic / (t_enc * 1e-3) / 1e9 / peak, "decode": algorithmic / (t_dec * 1e-3) / 1e9 / peak},
        "encode_ms": t_enc, "decode_ms": t_dec,
    }


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from charls_b200 import capi
    from charls_b200.batch import BatchCodec

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    lib = capi.default_library()
    lib.check(lib.charlsx_set_device(local_rank))
    w, h, bits, cc, near, ilv, xf = WORKLOADS[args.workload]
    F = args.frames
    frames = make_frames(torch, device, F, args.workload, first_seed=1234 + rank * F, content=args.content)
    if args.row_pitch:
        # the same frames with `row_pitch` bytes between their lines: views into wider tensors
        assert cc == 1 and args.row_pitch % frames.element_size() == 0 and args.row_pitch >= w * frames.element_size()
        wide = torch.zeros((F, h, args.row_pitch // frames.element_size()), device=device, dtype=frames.dtype)
        wide[:, :, :w] = frames
        frames = wide[:, :, :w]
    codec = BatchCodec(w, h, bits, cc, near_lossless=near, interleave_mode=ilv, color_transformation=xf,
                       restart_interval=args.restart_interval, offset_table=not args.no_offset_table, row_stride=args.row_pitch,
                       lib=lib)
    if args.content == "noise":
        codec.stream_capacity *= 2  # incompressible input expands (about 9.5 bits per 8-bit sample)
    streams = torch.empty((F, codec.stream_capacity), device=device, dtype=torch.uint8)
    if args.row_pitch:
        decoded = torch.zeros((F, h, args.row_pitch // frames.element_size()), device=device, dtype=frames.dtype)[:, :, :w]
    else:
        decoded = torch.empty_like(frames)
    raw_bytes = frames[0].numel() * frames.element_size()

    def step():
        sizes = codec.encode(frames, streams)
        t_enc = codec.last_coder_kernel_ms()
        codec.decode(streams, sizes, decoded)
        t_dec = codec.last_coder_kernel_ms()
        return sizes, t_enc, t_dec

    for _ in range(max(args.warmup, 3)):
        sizes, _, _ = step()
    torch.cuda.synchronize()

    # parity spot check outside the timed region: the round trip must reproduce the frames (within NEAR)
    if near == 0:
        assert torch.equal(decoded, frames), "round trip mismatch"
    else:
        a = decoded.to(torch.int32) & 0xFFFF if bits > 8 else decoded.to(torch.int32)
        b = frames.to(torch.int32) & 0xFFFF if bits > 8 else frames.to(torch.int32)
        assert int((a - b).abs().max()) <= near, "near-lossless bound violated"

    # one stream of the batch and its frame go to the host: the reference decodes it in the cpu_baseline leg (after the timed
    # regions), so that a bench line is never just our decoder agreeing with our encoder
    check_stream = streams[0, : sizes[0]].cpu().numpy().tobytes() if rank == 0 else None
    check_frame = frames[0].cpu().numpy() if rank == 0 else None

    launches_before = C.c_uint64()
    lib.check(lib.charlsx_get_kernel_launch_count(C.byref(launches_before)))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    enc_ms, dec_ms = [], []
    for _ in range(args.steps):
        sizes, t_enc, t_dec = step()
        enc_ms.append(t_enc)
        dec_ms.append(t_dec)
    stop.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches_after = C.c_uint64()
    lib.check(lib.charlsx_get_kernel_launch_count(C.byref(launches_after)))

    elapsed_ms = torch.tensor([start.elapsed_time(stop)], device=device, dtype=torch.float64)
    comp_total = torch.tensor([float(sum(sizes))], device=device, dtype=torch.float64)
    kernel_ms = torch.tensor([float(np.mean(enc_ms)), float(np.mean(dec_ms))], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kernel_ms, op=dist.ReduceOp.MAX)
        # control plane only: every rank learns every frame's stream size (the global offset table of the batch)
        gathered = [torch.empty(F, device=device, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gathered, torch.tensor(sizes, device=device, dtype=torch.int64))
        comp_total = torch.stack(gathered).sum().to(torch.float64).reshape(1)
    else:
        comp_total = comp_total

    ms_per_step = float(elapsed_ms.item()) / args.steps
    pixels_per_step = world * F * w * h
    value = pixels_per_step / (ms_per_step * 1e-3) / 1e6
    comp_per_frame = float(comp_total.item()) / (world * F)

    # ---- e2e through the C ABI with pinned host buffers (rank-local, then summed)
    e2e = None
    if not args.no_e2e:
        n = min(args.e2e_frames, F)
        frames_host = torch.empty((n,) + tuple(frames.shape[1:]), dtype=frames.dtype, pin_memory=True)
        frames_host.copy_(frames[:n])
        streams_host = torch.empty((n, codec.stream_capacity), dtype=torch.uint8, pin_memory=True)
        out_host = torch.empty_like(frames_host, pin_memory=True)
        torch.cuda.synchronize()
        passes = max(1, args.e2e_passes)
        reps = max(3, args.steps)
        cpus_per_rank = max(1, effective_cpus() // world)

        def measure(threads, inflight):
            for _ in range(2):
                e2e_round_trip(lib, frames_host, streams_host, out_host, args.workload, threads, not args.e2e_phases, 1, inflight)
            out_host.zero_()
            if world > 1:
                dist.barrier()
            total, sizes_ = 0.0, None
            for _ in range(reps):
                dt, sizes_ = e2e_round_trip(lib, frames_host, streams_host, out_host, args.workload, threads, not args.e2e_phases,
                                            passes, inflight)
                total += dt
            if near == 0:
                assert torch.equal(out_host, frames_host), "e2e round trip mismatch"
            t = torch.tensor([total / reps], device=device, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return world * passes * n * w * h / float(t.item()) / 1e6, sizes_

        # Headline: the reference's own calls (charls_jpegls_encoder_encode_from_buffer / charls_jpegls_decoder_decode_to_buffer,
        # synchronous), one image in flight per host thread.  The ranks of a box share its cores: callers mostly wait for the
        # GPU, so twice the cores are handed out, but not more (waiters that find no core slow everybody down).
        threads_one = max(1, min(args.e2e_one_part_threads, effective_cpus(), max(4, 2 * effective_cpus() // world)))
        value_one_part, e2e_sizes = measure(threads_one, 1)
        # Beside it: the two-part forms of the same calls (charlsx_*_begin / _end: "issue" and "complete"), several codec objects
        # in flight per host thread -- the same throughput from half as many threads
        threads = max(2, min(args.e2e_threads, cpus_per_rank))
        inflight = max(args.e2e_inflight, n // threads) if args.e2e_inflight > 1 else 1
        value_two_part, _ = measure(threads, inflight)
        comp = sum(e2e_sizes)

        # One image, one thread: what a single synchronous call costs (copy in, kernels, copy out)
        lat = {"enc": [], "dec": []}
        fi_one = capi.FrameInfo(w, h, bits, cc)
        for _ in range(12):
            e = lib.charls_jpegls_encoder_create()
            lib.check(lib.charls_jpegls_encoder_set_frame_info(e, C.byref(fi_one)))
            lib.check(lib.charls_jpegls_encoder_set_near_lossless(e, near))
            lib.check(lib.charls_jpegls_encoder_set_interleave_mode(e, ilv))
            lib.check(lib.charls_jpegls_encoder_set_color_transformation(e, xf))
            lib.check(lib.charls_jpegls_encoder_set_destination_buffer(e, streams_host[0].data_ptr(), streams_host.shape[1]))
            t0 = time.perf_counter()
            lib.check(lib.charls_jpegls_encoder_encode_from_buffer(e, frames_host[0].data_ptr(), raw_bytes, 0))
            lat["enc"].append(time.perf_counter() - t0)
            one_size = C.c_size_t()
            lib.check(lib.charls_jpegls_encoder_get_bytes_written(e, C.byref(one_size)))
            lib.charls_jpegls_encoder_destroy(e)
            d = lib.charls_jpegls_decoder_create()
            lib.check(lib.charls_jpegls_decoder_set_source_buffer(d, streams_host[0].data_ptr(), one_size.value))
            lib.check(lib.charls_jpegls_decoder_read_header(d))
            t0 = time.perf_counter()
            lib.check(lib.charls_jpegls_decoder_decode_to_buffer(d, out_host[0].data_ptr(), raw_bytes, 0))
            lat["dec"].append(time.perf_counter() - t0)
            lib.charls_jpegls_decoder_destroy(d)
        one_image_encode_ms = float(np.median(lat["enc"][2:])) * 1e3
        one_image_decode_ms = float(np.median(lat["dec"][2:])) * 1e3

        # What the box can copy: all ranks move the same bytes host->device and device->host at the same time (pinned buffers,
        # two streams, no kernel).  On the multi-GPU boxes of this pool the GPUs share the host's PCIe / memory path
        # (tools/pcie_probe_ranks.py: 96 GB/s both ways for one GPU, 104 for two, 120 for four), so this -- not the codec --
        # bounds e2e from two GPUs on.
        s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()
        scratch = torch.empty_like(frames[:n])
        copy_reps = 3

        def copy_both_ways():
            with torch.cuda.stream(s_up):
                scratch.copy_(frames_host, non_blocking=True)
            with torch.cuda.stream(s_down):
                out_host.copy_(frames[:n], non_blocking=True)

        copy_both_ways()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(copy_reps):
            copy_both_ways()
        s_up.synchronize()
        s_down.synchronize()
        c1.record()
        torch.cuda.synchronize()
        tc = torch.tensor([c0.elapsed_time(c1) / copy_reps / 1e3], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        copy_gbs_each_way = world * n * raw_bytes / float(tc.item()) / 1e9
        bytes_each_way_per_pixel = (n * raw_bytes + comp) / (n * w * h)
        copy_ceiling = copy_gbs_each_way * 1e9 / bytes_each_way_per_pixel / 1e6
        del scratch
        e2e = {
            "value": value_one_part, "unit": "MPixels/s",
            "h2d_bytes_per_step": world * passes * (n * raw_bytes + comp), "d2h_bytes_per_step": world * passes * (comp + n * raw_bytes),
            "frames_per_step": world * passes * n, "pinned_frame_buffers": world * n, "host_threads": threads_one,
            "objects_in_flight_per_thread": 1,
            "order": "all frames encoded, then all decoded" if args.e2e_phases else "each frame encoded, then decoded, by one worker",
            "api": "charls_jpegls_encoder_encode_from_buffer + charls_jpegls_decoder_decode_to_buffer, pinned host buffers",
            "two_part_calls_value": value_two_part, "two_part_calls_host_threads": threads, "two_part_calls_in_flight_per_thread": inflight,
            "two_part_calls_api": "charlsx_jpegls_encoder_encode_from_buffer_begin/_end + charlsx_jpegls_decoder_decode_to_buffer_begin/_end",
            "one_image_encode_call_ms": one_image_encode_ms, "one_image_decode_call_ms": one_image_decode_ms,
            "box_copy_gbs_each_way": copy_gbs_each_way, "box_copy_ceiling_mpix_s": copy_ceiling,
            "frac_of_box_copy_ceiling": value_one_part / copy_ceiling,
            "box_copy_ceiling_how": "all ranks copy the same pinned buffers up and down at the same time, no kernel (CUDA events, max over ranks)",
        }

        # ---- the same host buffers through the host-batch extension: one call per direction from one host thread, the
        # library stages chunks of frames and overlaps their copies with the kernels
        pixels_in = [frames_host[i % n] for i in range(passes * n)]
        streams_io = [streams_host[i % n] for i in range(passes * n)]
        pixels_out = [out_host[i % n] for i in range(passes * n)]
        t_batch = 0.0
        batch_reps = max(3, reps // 2)
        for rep in range(2 + batch_reps):
            out_host.zero_()
            if world > 1 and rep == 2:
                dist.barrier()
            t0 = time.perf_counter()
            batch_sizes = codec.encode_host(pixels_in, streams_io)
            codec.decode_host(streams_io, batch_sizes, pixels_out)
            if rep >= 2:
                t_batch += time.perf_counter() - t0
        if near == 0:
            assert torch.equal(out_host, frames_host), "host-batch round trip mismatch"
        tb = torch.tensor([t_batch / batch_reps], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tb, op=dist.ReduceOp.MAX)
        e2e["host_batch_value"] = world * passes * n * w * h / float(tb.item()) / 1e6
        e2e["host_batch_api"] = "charlsx_batch_encode_host, then charlsx_batch_decode_host (extension), one host thread per GPU, same pinned buffers"

        # ---- both PCIe directions at once: one thread encodes half of the frames while a second one decodes the streams of the
        # other half (two batch objects), then the halves swap roles; a round trip still is one encode plus one decode of a frame
        if n >= 2:
            codec_b = BatchCodec(w, h, bits, cc, near_lossless=near, interleave_mode=ilv, color_transformation=xf,
                                 restart_interval=args.restart_interval, offset_table=not args.no_offset_table, lib=lib)
            lower = [i % (n // 2) for i in range(passes * (n // 2))]            # buffers 0 .. n/2-1, each `passes` times
            upper = [n // 2 + i % (n - n // 2) for i in range(passes * (n - n // 2))]  # buffers n/2 .. n-1: the halves share nothing
            first = ([frames_host[i] for i in lower], [streams_host[i] for i in lower], [out_host[i] for i in lower], [batch_sizes[i] for i in lower])
            second = ([frames_host[i] for i in upper], [streams_host[i] for i in upper], [out_host[i] for i in upper], [batch_sizes[i] for i in upper])

            def two_way(encode_side, decode_side):
                worker = threading.Thread(target=lambda: codec_b.decode_host(decode_side[1], decode_side[3], decode_side[2]))
                worker.start()
                codec.encode_host(encode_side[0], encode_side[1])
                worker.join()

            t_two = 0.0
            for rep in range(1 + batch_reps):
                if world > 1 and rep == 1:
                    dist.barrier()
                t0 = time.perf_counter()
                two_way(first, second)  # streams of `second` are from the previous encode, `first` gets new ones
                two_way(second, first)
                if rep >= 1:
                    t_two += time.perf_counter() - t0
            if near == 0:
                assert torch.equal(out_host, frames_host), "two-way host-batch mismatch"
            tt = torch.tensor([t_two / batch_reps], device=device, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e["host_batch_two_way_value"] = world * passes * n * w * h / float(tt.item()) / 1e6
            e2e["host_batch_two_way_api"] = "charlsx_batch_encode_host of one half of the frames while a second thread runs charlsx_batch_decode_host on the other half's streams"
            codec_b.close()

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_kind = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    else:
        peak, peak_kind = 6650.0, "fallback (B200_PROFILING.md)"

    # ---- the metric's other inputs (16-bit RGB and the 12-bit near-lossless frame beside the 8-bit one): measured like `value`
    also = {}
    if args.also != "none" and args.content == "smooth":
        others = [wl for wl in ("cfg4", "cfg3", "cfg2") if wl != args.workload][:2] if args.also == "auto" else [args.also]
        del frames, streams, decoded
        torch.cuda.empty_cache()
        for other in others:
            also[other] = secondary_workload(torch, dist, lib, device, rank, world, other, args.frames, max(3, args.steps // 4), peak,
                                             not args.no_offset_table)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the entropy-coding kernels (algorithmic bytes = raw + compressed)
    algorithmic = F * (raw_bytes + comp_per_frame)
    t_enc, t_dec = float(kernel_ms[0].item()), float(kernel_ms[1].item())
    # DRAM traffic per launch comes from an ncu capture (profiles/ncu_traffic.json, written by tools/ncu_traffic.py together
    # with a hash of the kernel sources and the commit it was taken at); it is null when the kernels have changed since.
    traffic_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
    traffic_fresh = traffic.get("kernels_sha") == kernel_sources_sha()
    traffic_note = (f"ncu capture at {traffic.get('git', '?')} (profiles/ncu_traffic.json)" if traffic_fresh
                    else f"null: kernel sources changed since the capture at {traffic.get('git', '?')}")

    def roof(name, ms):
        achieved = algorithmic / (ms * 1e-3) / 1e9
        return {"kernel": name, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic.get("entries", {}).get(f"{name}:{args.workload}:{F}") if traffic_fresh else None,
                "traffic_source": traffic_note, "ms_per_launch": ms, "peak_source": peak_kind,
                "algorithmic_bytes_per_launch": algorithmic}

    fast = args.restart_interval == 1
    # the per-lane kernels without shared-memory tiles: line interleave only (device frames whose rows do not start on word
    # boundaries are copied to an aligned pitch around the tile kernels, Engine::repitch: in `value`, not in the kernel's time)
    tiled = ilv != 1
    family = ("tiled" if tiled else "fast") if fast else "general"
    roofs = {"encode": roof(f"k_encode_{family}", t_enc), "decode": roof(f"k_decode_{family}", t_dec)}
    dominant = dict(roofs["encode"] if t_enc >= t_dec else roofs["decode"])
    other_kernel = roofs["decode"] if t_enc >= t_dec else roofs["encode"]
    dominant["other_kernel"] = other_kernel["kernel"]
    dominant["other_kernel_frac"] = other_kernel["frac"]
    dominant["other_kernel_ms_per_launch"] = other_kernel["ms_per_launch"]
    config_also = {}
    for name, r in also.items():
        # flat scalars: the driver's record keeps the scalar members of `config` and `roofline`
        config_also[f"{name}_workload"] = r["workload"]
        config_also[f"{name}_value_mpix_s"] = r["value"]
        config_also[f"{name}_encode_mpix_s"] = r["encode_mpix_s"]
        config_also[f"{name}_decode_mpix_s"] = r["decode_mpix_s"]
        config_also[f"{name}_ratio"] = r["ratio"]
        dominant[f"{name}_encode_frac"] = r["roofline_frac"]["encode"]
        dominant[f"{name}_decode_frac"] = r["roofline_frac"]["decode"]
        dominant[f"{name}_encode_ms_per_launch"] = r["encode_ms"]
        dominant[f"{name}_decode_ms_per_launch"] = r["decode_ms"]

    cpu_baseline = None
    if world == 1 and not args.no_cpu and os.path.exists(REF_LIB):
        ref = capi.CharlsLibrary(REF_LIB, extensions=False)
        threads = effective_cpus()
        hf = host_frames(args.workload, min(threads, 4))
        hf = [hf[i % len(hf)] for i in range(threads)]
        buffers = {}
        cpu_reference_step(ref, hf, args.workload, threads, buffers)  # warm
        reps, t = 0, 0.0
        while t < 8.0 and reps < 20:
            t += cpu_reference_step(ref, hf, args.workload, threads, buffers)
            reps += 1
        cpu_baseline = {"value": threads * reps * w * h / t / 1e6, "unit": "MPixels/s", "cores": threads, "kind": "reference",
                        "sample": f"{threads} threads x {reps} frames of {args.workload}: unmodified reference encode (no restart markers) + decode"}
        # the reference as the checker of the GPU arm: it decodes frame 0's stream of the timed batch (side table and all)
        from charls_b200 import codec as host_codec

        ref_pixels, _, _ = host_codec.decode(check_stream, lib=ref)
        mine = check_frame.view(np.uint16) if bits > 8 else check_frame
        if near == 0:
            agrees = bool(np.array_equal(ref_pixels.reshape(mine.shape), mine))
        else:
            agrees = int(np.abs(ref_pixels.reshape(mine.shape).astype(np.int64) - mine.astype(np.int64)).max()) <= near
        assert agrees, "the reference decodes a stream of the timed batch to something else than its frame"
        cpu_baseline["reference_decodes_bench_stream"] = agrees

    out = {
        "metric": METRIC, "value": value, "unit": "MPixels/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8" if bits <= 8 else "u16", "data": "synthetic",
        "config": {"workload": workload_name(args), "frames_per_step": world * F, "restart_interval": args.restart_interval,
                   "compressed_bytes_per_frame": comp_per_frame, "ratio": raw_bytes / comp_per_frame,
                   "cache": f"inputs larger than L2: {F * raw_bytes / 1e6:.0f} MB raw + {F * comp_per_frame / 1e6:.0f} MB streams per GPU vs 126 MB L2",
                   "sharding": "frames split across ranks, no data-path collective; NCCL all_gather of stream sizes only",
                   "offset_table": not args.no_offset_table and args.restart_interval != 0,
                   **config_also},
        "encode_mpix_s": world * F * w * h / (t_enc * 1e-3) / 1e6, "decode_mpix_s": world * F * w * h / (t_dec * 1e-3) / 1e6,
        "roofline": dominant, "roofline_all": roofs, "cpu_baseline": cpu_baseline, "e2e": e2e, "also": also or None, "clocks": clocks,
        "gpu_launches": int(launches_after.value - launches_before.value),
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
