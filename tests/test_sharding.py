"""Multi-GPU plumbing on CPU: world_size-2 gloo processes shard a batch, code their frames with the kernel code
(host emulation -- the ranks here have no GPU) and all-gather the stream sizes."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

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
