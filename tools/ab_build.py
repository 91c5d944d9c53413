"""Builds variants of the library for A/B kernel experiments on the GPU box.
usage: python tools/ab_build.py [--subset] name1:-DJLS_X=1,-DJLS_Y=0 name2:-DJLS_X=0 ...
Each variant lands in charls_b200/build/variants/<name>/libcharls.so.3 (travels with gpurun, ignored by git); run the
bench against one with CHARLS_B200_LIBRARY=<path> python bench.py ...   Variants compile in parallel.
--subset adds -DJLS_DEV_SUBSET (one- and three-component kernels only: half the compile time; enough for cfg2/3/4)."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from charls_b200 import build as B  # noqa: E402

args = [a for a in sys.argv[1:] if a != "--subset"]
subset = ["-DJLS_DEV_SUBSET"] if "--subset" in sys.argv else []
B.build()
jobs = []
for spec in args:
    name, _, flags = spec.partition(":")
    out_dir = os.path.join(B.OBJ_DIR, "variants", name)
    os.makedirs(out_dir, exist_ok=True)
    obj = os.path.join(out_dir, "jls_kernels.cu.o")
    cmd = [B.NVCC, *B.ARCH, *B.COMMON, *subset, *[f for f in flags.split(",") if f], "-x", "cu", "-c",
           os.path.join(B.CSRC, "jls_kernels.cu"), "-o", obj]
    jobs.append((name, out_dir, obj, subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)))
for name, out_dir, obj, proc in jobs:
    _, err = proc.communicate()
    if proc.returncode != 0:
        sys.stderr.write(err)
        raise SystemExit(f"variant {name} failed to compile")
    objects = [obj] + [os.path.join(B.OBJ_DIR, s.replace("/", "_") + ".o") for s in B.SOURCES if s != "jls_kernels.cu"]
    lib = os.path.join(out_dir, "libcharls.so.3")
    subprocess.check_call([B.NVCC, *B.ARCH, "-shared", "-o", lib, *objects, "-Xlinker", "-soname,libcharls.so.3", "-Xlinker",
                           "--version-script=" + os.path.join(B.CSRC, "exports.map"), "-cudart", "static", "-Xcompiler",
                           "-static-libstdc++,-static-libgcc"])
    print(lib)
