"""One large frame across the GPUs of a box, everything on the devices (charls_b200.sharding.encode_frame_split_device /
decode_frame_split_device): strips coded per rank, entropy-coded bytes exchanged with NCCL, joined on the device.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tools/split_frame_check.py [H W]

Checks (rank 0): the joined stream is byte for byte what ONE GPU writes for the whole frame; the reference (oracle/_ref, if
built) decodes it to the frame; the split decode -- every rank fetches only its strip through the side table of interval
offsets -- returns the frame.  Prints one line and exits non-zero on any difference.  Works with one process too."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from charls_b200 import capi, codec, sharding
from charls_b200.batch import BatchCodec


def main():
    h, w = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16384, 16384)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib = capi.default_library()
    lib.check(lib.charlsx_set_device(local))

    x = torch.arange(w, device=device, dtype=torch.float32)[None, :]
    y = torch.arange(h, device=device, dtype=torch.float32)[:, None]
    gen = torch.Generator(device=device)
    gen.manual_seed(77)
    frame = torch.clamp(204 * (0.5 + 0.25 * torch.sin(x / 97.0) + 0.25 * torch.cos(y / 131.0)) +
                        torch.randn((h, w), device=device, generator=gen) * 2.55, 0, 255).to(torch.uint8)

    def strip_codec(lines, table=False):
        return BatchCodec(w, lines, 8, 1, restart_interval=1, offset_table=table, lib=lib)

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    stream, size = sharding.encode_frame_split_device(frame, strip_codec, dist if world > 1 else None)
    torch.cuda.synchronize()
    t_split = time.perf_counter() - t0

    ok = True
    if rank == 0:
        whole = strip_codec(h)
        single = torch.empty((1, whole.stream_capacity), dtype=torch.uint8, device=device)
        (single_size,) = whole.encode(frame.unsqueeze(0), single)
        same = single_size == size and torch.equal(single[0, :size], stream)
        ok = ok and same
        reference = "not built"
        ref_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libcharls_ref.so")
        if os.path.exists(ref_path):
            ref = capi.CharlsLibrary(ref_path, extensions=False)
            px, _, _ = codec.decode(stream.cpu().numpy().tobytes(), lib=ref)
            reference = "decodes it to the frame" if np.array_equal(px, frame.cpu().numpy()) else "DIFFERS"
            ok = ok and reference != "DIFFERS"
        whole.close()
    # decode: the whole-frame stream WITH the side table (written by one encoder), split by table
    tabled_codec = strip_codec(h, table=True)
    tabled = torch.empty((1, tabled_codec.stream_capacity), dtype=torch.uint8, device=device)
    (tabled_size,) = tabled_codec.encode(frame.unsqueeze(0), tabled)
    tabled_codec.close()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    decoded = sharding.decode_frame_split_device(tabled[0], tabled_size, strip_codec, dist if world > 1 else None)
    torch.cuda.synchronize()
    t_decode = time.perf_counter() - t0
    decoded_ok = torch.equal(decoded, frame)
    ok = ok and decoded_ok
    if rank == 0:
        print(f"split_frame_check: {world} rank(s), {w}x{h}: joined stream {size} bytes, identical to the single-GPU stream: {same}; "
              f"reference: {reference}; split decode through the offset table returns the frame: {decoded_ok}; "
              f"encode+exchange {t_split * 1e3:.1f} ms, decode+exchange {t_decode * 1e3:.1f} ms")
    if world > 1:
        flag = torch.tensor([1 if ok else 0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = bool(flag.item())
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


main()
