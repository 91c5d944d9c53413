"""Builds variants of the library for A/B kernel experiments on the GPU box.
usage: python tools/ab_build.py name1:-DJLS_X=1,-DJLS_Y=0 name2:-DJLS_X=0 ...
Each variant lands in charls_b200/build/variants/<name>/libcharls.so.3 (travels with gpurun, ignored by git); run the
bench against one with CHARLS_B200_LIBRARY=<path> python bench.py ..."""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from charls_b200 import build as B  # noqa: E402

B.build()
for spec in sys.argv[1:]:
    name, _, flags = spec.partition(":")
    out_dir = os.path.join(B.OBJ_DIR, "variants", name)
    os.makedirs(out_dir, exist_ok=True)
    obj = os.path.join(out_dir, "jls_kernels.cu.o")
    subprocess.check_call([B.NVCC, *B.ARCH, *B.COMMON, *[f for f in flags.split(",") if f], "-x", "cu", "-c",
                           os.path.join(B.CSRC, "jls_kernels.cu"), "-o", obj])
    objects = [obj] + [os.path.join(B.OBJ_DIR, s.replace("/", "_") + ".o") for s in B.SOURCES if s != "jls_kernels.cu"]
    lib = os.path.join(out_dir, "libcharls.so.3")
    subprocess.check_call([B.NVCC, *B.ARCH, "-shared", "-o", lib, *objects, "-Xlinker", "-soname,libcharls.so.3", "-Xlinker",
                           "--version-script=" + os.path.join(B.CSRC, "exports.map"), "-cudart", "static", "-Xcompiler",
                           "-static-libstdc++,-static-libgcc"])
    print(lib)
