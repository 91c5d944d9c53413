"""Builds the in-tree shared library charls_b200/lib/libcharls.so.3 (sm_100a only) with nvcc.

Usage: python -m charls_b200.build [--force]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIB_DIR, "libcharls.so.3")
OBJ_DIR = os.path.join(HERE, "build")

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wextra", "-DCHARLS_B200_BUILD",
          "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]

SOURCES = [
    "jls_kernels.cu",
    "engine.cu",
    "host/stream_reader.cpp",
    "host/encoder.cpp",
    "host/decoder.cpp",
    "host/misc.cpp",
    "host/batch.cpp",
]


def _headers():
    out = [os.path.join(ROOT, "include", "charls_b200.h")]
    for d, _, files in os.walk(CSRC):
        out += [os.path.join(d, f) for f in files if f.endswith((".h", ".hpp", ".cuh"))]
    return out


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = _headers()
    objects = []
    procs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ_DIR, src.replace("/", "_") + ".o")
        objects.append(obj)
        if force or _stale(obj, [path] + headers):
            cmd = [NVCC, *ARCH, *COMMON, "-x", "cu", "-c", path, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, proc in procs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            failed = True
            sys.stderr.write(f"--- nvcc failed for {src}\n{out}\n")
        elif verbose and out.strip():
            print(out)
    if failed:
        raise RuntimeError("charls_b200: compilation failed")
    if force or procs or _stale(LIB, objects + [os.path.join(CSRC, "exports.map")]):
        cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objects, "-Xlinker", "-soname,libcharls.so.3", "-Xlinker", "--version-script=" + os.path.join(CSRC, "exports.map"), "-cudart", "static",
               "-Xcompiler", "-static-libstdc++,-static-libgcc"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    build_driver(force, verbose)
    return LIB


DRIVER = os.path.join(LIB_DIR, "libabi_driver.so")


def build_driver(force: bool = False, verbose: bool = False) -> str:
    """charls_b200/lib/libabi_driver.so: plain C++ threads that call a CharLS-compatible C ABI (csrc/driver)."""
    src = os.path.join(CSRC, "driver", "abi_driver.cpp")
    if force or _stale(DRIVER, [src]):
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall", "-Wextra", "-o", DRIVER, src, "-ldl",
               "-lpthread", "-static-libstdc++", "-static-libgcc"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return DRIVER


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
