"""Small workload for profiling the general path (one CUDA thread per restart interval): 64 frames of 512 x 512 8-bit S_smooth
with no restart interval (what the reference writes), encoded and decoded through the batch interface.
usage: ncu --set full --import-source on -k regex:general -c 2 python tools/prof_general.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import WORKLOADS, make_frames
from charls_b200 import capi
from charls_b200.batch import BatchCodec

WORKLOADS["small"] = (512, 512, 8, 1, 0, 0, 0)
device = torch.device("cuda", 0)
frames = make_frames(torch, device, 64, "small", 1234)
codec = BatchCodec(512, 512, 8, 1, restart_interval=0, lib=capi.default_library())
streams = torch.empty((64, codec.stream_capacity), device=device, dtype=torch.uint8)
out = torch.empty_like(frames)
for _ in range(2):
    sizes = codec.encode(frames, streams)
    t_enc = codec.last_coder_kernel_ms()
    codec.decode(streams, sizes, out)
    t_dec = codec.last_coder_kernel_ms()
torch.cuda.synchronize()
assert torch.equal(out, frames)
px = 64 * 512 * 512
print(f"general path, 64 x 512x512: encode {t_enc:.1f} ms ({px / t_enc / 1e3:.1f} MPix/s), decode {t_dec:.1f} ms ({px / t_dec / 1e3:.1f} MPix/s)")
