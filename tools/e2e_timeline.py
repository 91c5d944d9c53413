"""Reads a CHARLS_B200_TRACE file (charls_b200/csrc/engine.cu: Trace) and says where the single-image calls spend their time.
usage: python tools/e2e_timeline.py trace.txt [skip_fraction]
Per call: host_ms = entered, enqueued, outcome known, done; gpu_ms = stream reached the call, input copy done, kernels done,
output copy done (CUDA events against one base event)."""
import sys

import numpy as np

rows = np.loadtxt(sys.argv[1], ndmin=2)
skip = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
rows = rows[int(len(rows) * skip):]  # steady state: drop warm-up rounds
for kind, name in ((0, "encode"), (1, "decode")):
    r = rows[rows[:, 1] == kind]
    if not len(r):
        continue
    h, g = r[:, 2:6], r[:, 6:10]
    print(f"{name}: {len(r)} calls")
    print(f"  host: enqueue {np.mean(h[:, 1] - h[:, 0]):6.3f} ms | wait for outcome {np.mean(h[:, 2] - h[:, 1]):6.3f} | "
          f"output copy + wait {np.mean(h[:, 3] - h[:, 2]):6.3f} | call {np.mean(h[:, 3] - h[:, 0]):6.3f}")
    print(f"  gpu : input copy (queue + transfer) {np.mean(g[:, 1] - g[:, 0]):6.3f} ms | kernels (queue + run) "
          f"{np.mean(g[:, 2] - g[:, 1]):6.3f} | host turn-around + output copy {np.mean(g[:, 3] - g[:, 2]):6.3f} | "
          f"total {np.mean(g[:, 3] - g[:, 0]):6.3f}")
g_all = rows[:, 6:10]
span = g_all[:, 3].max() - g_all[:, 0].min()
print(f"all: {len(rows)} calls in {span:.1f} ms of GPU timeline = {len(rows) / 2 / span * 1e3:.0f} round trips/s")
