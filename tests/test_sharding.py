"""Multi-GPU plumbing on CPU: world_size-2 gloo processes shard a batch, code their frames with the kernel code
(host emulation -- the ranks here have no GPU) and all-gather the stream sizes."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from charls_b200 import sharding
from charls_b200.sharding import frame_range, gather_sizes, offset_table


def test_frame_range_partitions_everything():
    for total in (1, 7, 8, 1024, 1025):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                seen += list(frame_range(total, world, r))
            assert seen == list(range(total))
            lengths = [len(frame_range(total, world, r)) for r in range(world)]
            assert max(lengths) - min(lengths) <= 1


def _worker(rank, world, port, total, result_dir):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from tests.hostemu_lib import HostEmu
    from tests.support import oracle, s_smooth

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    o, he = oracle(), HostEmu()
    mine = frame_range(total, world, rank)
    sizes, streams = [], {}
    for i in mine:
        img = s_smooth(8, 40, 8, seed=1234 + i)
        sp = o.params(40, 8, 8, 1, 0, 0, 0, None, 1)
        n, data = he.encode(he.params(sp), img, 4096)
        assert n > 0
        sizes.append(n)
        streams[i] = data
    all_sizes = gather_sizes(sizes, total, dist)
    offsets, total_bytes = offset_table(all_sizes)
    np.save(os.path.join(result_dir, f"rank{rank}.npy"), np.array(all_sizes + offsets + [total_bytes], dtype=np.int64))
    # every rank's own frames agree with the oracle (what a single process would have produced)
    for i in mine:
        img = s_smooth(8, 40, 8, seed=1234 + i)
        assert streams[i] == o.encode_scan(o.params(40, 8, 8, 1, 0, 0, 0, None, 1), img)
        assert all_sizes[i] == len(streams[i])
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    total = 7  # uneven on purpose
    mp.spawn(_worker, args=(2, port, total, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "rank0.npy")
    b = np.load(tmp_path / "rank1.npy")
    assert np.array_equal(a, b) and len(a) == 2 * total + 1
    assert a[-1] == a[:total].sum()


# -- one frame across several ranks ---------------------------------------------------------------------

STRIP_CASES = ((37, 50, 8, 1, 0, 0), (64, 33, 12, 1, 0, 2), (21, 40, 16, 3, 2, 0), (8, 9, 8, 1, 0, 0), (5, 9, 8, 1, 0, 0),
               (40, 31, 8, 3, 0, 0), (19, 20, 16, 4, 0, 1))  # the last two: planar frames, one scan per component


def _strip_image(h, w, bits, cc, ilv=2):
    from tests.support import s_mixed

    return s_mixed(h, w, bits, cc, layout="planar" if ilv == 0 else "interleaved") if cc > 1 else s_mixed(h, w, bits)


def _line_axis(cc, ilv):
    return 1 if cc > 1 and ilv == 0 else 0


def _lines(image, lines, axis):
    return np.ascontiguousarray(np.take(image, lines, axis=axis))


def test_strip_ranges_cover_the_frame_on_multiples_of_eight():
    for height in (1, 5, 8, 9, 37, 64, 4096, 65535):
        for world in (1, 2, 3, 8):
            ranges = [sharding.strip_range(height, world, r) for r in range(world)]
            assert [x for r in ranges for x in r] == list(range(height))
            assert all(r.start % 8 == 0 for r in ranges if len(r))


def test_strips_stitch_to_the_single_encoder_stream(oracle, reference):
    """Strips coded on their own and joined by the restart marker between them are the stream one encoder writes for the
    whole frame, byte for byte (headers included); the unmodified reference decodes it; cutting it again gives the strips."""
    from charls_b200 import codec

    for h, w, bits, cc, ilv, near in STRIP_CASES:
        image = _strip_image(h, w, bits, cc, ilv)
        axis = _line_axis(cc, ilv)
        whole = oracle.encode_image(image, bits, near=near, ilv=ilv, ri=1)
        expected, _ = oracle.decode_image(whole)
        for world in (1, 2, 3, 8):
            ranges = [sharding.strip_range(h, world, r) for r in range(world)]
            strips = [oracle.encode_image(_lines(image, r, axis), bits, near=near, ilv=ilv, ri=1) if len(r) else b"" for r in ranges]
            stitched = sharding.stitch_strips(strips, [len(r) for r in ranges])
            assert stitched == whole, (h, w, bits, world)
            pixels, _, _ = codec.decode(stitched, lib=reference)
            assert np.array_equal(pixels, expected)
            parts = sharding.split_stream(whole, world)
            assert [p if p is not None else b"" for p in parts] == strips, (h, w, bits, world)


def _strip_worker(rank, world, port, result_dir):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from tests.support import oracle

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    o = oracle()
    for case, (h, w, bits, cc, ilv, near) in enumerate(STRIP_CASES):
        image = _strip_image(h, w, bits, cc, ilv)
        axis = _line_axis(cc, ilv)
        stream = sharding.encode_frame_split(image, lambda rows: o.encode_image(rows, bits, near=near, ilv=ilv, ri=1), dist, axis)
        assert stream == o.encode_image(image, bits, near=near, ilv=ilv, ri=1)
        pixels = sharding.decode_frame_split(stream, lambda s: o.decode_image(s)[0], dist, axis)
        assert np.array_equal(pixels, o.decode_image(stream)[0])
        if rank == 0:
            with open(os.path.join(result_dir, f"case{case}.jls"), "wb") as f:
                f.write(stream)
    dist.barrier()
    dist.destroy_process_group()


def test_one_frame_on_two_ranks_gloo(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_strip_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert len(list(tmp_path.iterdir())) == len(STRIP_CASES)


def test_strip_helpers_reject_what_they_cannot_join(oracle, reference):
    from charls_b200 import codec

    image = _strip_image(24, 16, 8, 1)
    whole = oracle.encode_image(image, 8, ri=1)
    first = oracle.encode_image(image[:12], 8, ri=1)
    second = oracle.encode_image(image[12:], 8, ri=1)
    with pytest.raises(ValueError):
        sharding.stitch_strips([first, second], [12, 12])  # the second strip does not start on a multiple of eight lines
    with pytest.raises(ValueError):
        sharding.split_stream(oracle.encode_image(image, 8, ri=0), 2)  # no restart markers to cut at
    with pytest.raises(ValueError):
        sharding.stitch_strips([b"", b""], [0, 0])
    # fill bytes in front of a restart marker (ISO/IEC 10918-1 B.1.1.2) stay with the strip in front of them
    header, [(sos, payload)] = sharding._scans(whole)
    cut = payload.index(b"\xff\xd7")
    padded = header + sos + payload[:cut] + b"\xff" + payload[cut:] + b"\xff\xd9"
    parts = sharding.split_stream(padded, 3)
    pixels = np.concatenate([codec.decode(p, lib=reference)[0] for p in parts if p is not None], axis=0)
    assert np.array_equal(pixels, image)
    assert np.array_equal(codec.decode(padded, lib=reference)[0], image)
