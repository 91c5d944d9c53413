"""What the e2e path could reach if the codec cost nothing: the copies of a cfg2 round trip (frame up, stream down, stream up,
frame down; a stream synchronisation after each, as the synchronous C ABI has) from T host threads with one CUDA stream each,
pinned buffers, NO kernels - against the same bytes as four big copies in two directions at once (bench.py's copy ceiling).

    python tools/copy_pattern_probe.py [threads ...]        (one GPU; prints one line per thread count)
"""
import sys
import threading
import time

import torch

W = H = 4096
RAW = W * H
COMP = RAW // 2  # S_smooth compresses 2:1
FRAMES = 128


def main():
    dev = torch.device("cuda:0")
    thread_counts = [int(a) for a in sys.argv[1:]] or [4, 8, 16, 32]
    n_host = 32
    raw_host = torch.empty((n_host, RAW), dtype=torch.uint8, pin_memory=True)
    comp_host = torch.empty((n_host, COMP), dtype=torch.uint8, pin_memory=True)
    raw_dev = torch.empty((n_host, RAW), dtype=torch.uint8, device=dev)
    comp_dev = torch.empty((n_host, COMP), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()

    # the ceiling: everything at once, two streams
    s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()

    def big():
        with torch.cuda.stream(s_up):
            raw_dev.copy_(raw_host, non_blocking=True)
            comp_dev.copy_(comp_host, non_blocking=True)
        with torch.cuda.stream(s_down):
            comp_host.copy_(comp_dev, non_blocking=True)
            raw_host.copy_(raw_dev, non_blocking=True)

    big()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = FRAMES // n_host
    for _ in range(reps):
        big()
    torch.cuda.synchronize()
    t_big = time.perf_counter() - t0
    print(f"big copies, two streams: {FRAMES * (RAW + COMP) / t_big / 1e9:.1f} GB/s each way = {FRAMES * RAW / t_big / 1e9:.2f} GPix/s")

    for threads in thread_counts:
        def worker(t):
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                for f in range(t, FRAMES, threads):
                    i = f % n_host
                    raw_dev[i].copy_(raw_host[i], non_blocking=True)
                    stream.synchronize()
                    comp_host[i].copy_(comp_dev[i], non_blocking=True)
                    stream.synchronize()
                    comp_dev[i].copy_(comp_host[i], non_blocking=True)
                    stream.synchronize()
                    raw_host[i].copy_(raw_dev[i], non_blocking=True)
                    stream.synchronize()

        best = None
        for _ in range(3):
            pool = [threading.Thread(target=worker, args=(t,)) for t in range(threads)]
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for p in pool:
                p.start()
            for p in pool:
                p.join()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        print(f"{threads:3d} threads, per-image copies with syncs: {FRAMES * (RAW + COMP) / best / 1e9:.1f} GB/s each way = "
              f"{FRAMES * RAW / best / 1e9:.2f} GPix/s ({best / t_big:.2f} x the big copies' time)")


if __name__ == "__main__":
    main()
