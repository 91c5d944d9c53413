import sys, time, os
sys.path.insert(0, "/root/repo")
import torch
import bench
from charls_b200.batch import BatchCodec
w = h = 4096; n = 64
frames = bench.make_frames(torch, torch.device("cuda"), n, "cfg2", 1234)
frames_host = frames.cpu().pin_memory()
hb = BatchCodec(w, h, 8)
cap = hb.stream_capacity
streams_host = torch.empty((n, cap), dtype=torch.uint8).pin_memory()
out_host = torch.empty_like(frames_host).pin_memory()
F = [frames_host[i] for i in range(n)]; S = [streams_host[i] for i in range(n)]; O = [out_host[i] for i in range(n)]
for rep in range(3):
    print("--- rep", rep, file=sys.stderr, flush=True)
    t0 = time.perf_counter(); sizes = hb.encode_host(F, S); t1 = time.perf_counter(); hb.decode_host(S, sizes, O); t2 = time.perf_counter()
    print(f"rep {rep}: encode {(t1-t0)*1e3:.1f} ms decode {(t2-t1)*1e3:.1f} ms", file=sys.stderr, flush=True)
assert torch.equal(out_host, frames_host)
