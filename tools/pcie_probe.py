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

# ---- the copy pattern of the end-to-end bench without any kernel: 16 streams, per round trip 16.8 MB up, 8.4 MB down,
# 8.4 MB up, 16.8 MB down (cfg2 frame and its stream), all asynchronous
frame, comp = 4096 * 4096, 4096 * 4096 // 2
streams = [torch.cuda.Stream() for _ in range(16)]
hosts = [torch.empty(frame, dtype=torch.uint8).pin_memory() for _ in range(16)]
devs = [torch.empty(frame, dtype=torch.uint8, device="cuda") for _ in range(16)]


def pattern(sync_each):
    for k, s in enumerate(streams):
        with torch.cuda.stream(s):
            devs[k].copy_(hosts[k], non_blocking=True)
            hosts[k][:comp].copy_(devs[k][:comp], non_blocking=True)
            if sync_each:
                s.synchronize()
            devs[k][:comp].copy_(hosts[k][:comp], non_blocking=True)
            hosts[k].copy_(devs[k], non_blocking=True)


for sync_each in (False,):
    t = timed(lambda: pattern(sync_each), reps=4)
    per_direction = 16 * (frame + comp) / t / 1e9
    print(json.dumps({"pattern": "16 streams x (16.8 MB up, 8.4 MB down, 8.4 MB up, 16.8 MB down)", "each_direction_gbs": per_direction,
                      "round_trips_per_s": 16 / t}))
