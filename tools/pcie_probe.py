"""Host<->device copy bandwidth of the box with pinned buffers (context for bench.py's e2e number).
Prints one JSON line: GB/s host->device, device->host, and both directions at the same time on two streams."""
import json

import torch

n = 1 << 30
host_a = torch.empty(n, dtype=torch.uint8).pin_memory()
host_b = torch.empty(n, dtype=torch.uint8).pin_memory()
dev_a = torch.empty(n, dtype=torch.uint8, device="cuda")
dev_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / 1e3


def h2d():
    with torch.cuda.stream(s1):
        dev_a.copy_(host_a, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        host_b.copy_(dev_b, non_blocking=True)


def both():
    h2d()
    d2h()


t_h2d, t_d2h, t_both = timed(h2d), timed(d2h), timed(both)
print(json.dumps({"h2d_gbs": n / t_h2d / 1e9, "d2h_gbs": n / t_d2h / 1e9, "bidirectional_sum_gbs": 2 * n / t_both / 1e9}))
