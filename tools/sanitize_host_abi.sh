#!/bin/bash
# The host side of the ABI (charls_b200/csrc/host/*.cpp: state machines, JPEG-LS container reader / writer) compiled with
# UBSan and with ASan, linked with the regular CUDA objects, and driven by the ABI tests and the long differential walks
# (tools/parity_hunt.py abi) in the GPU-less container.  Needs charls_b200/build/*.o (python -c "import __graft_entry__ as g; g.build()").
# End of round 1: no finding with either sanitizer.
set -e
cd "$(dirname "$0")/.."
ROOT=$PWD
OUT=${1:-/tmp/charls_b200_sanitize}
for kind in undefined address; do
  mkdir -p $OUT/$kind
  for f in stream_reader encoder decoder misc batch; do
    g++ -std=c++17 -O1 -g -fPIC -fvisibility=hidden -DCHARLS_B200_BUILD -fsanitize=$kind -I/usr/local/cuda/include \
        -I$ROOT/charls_b200/csrc -I$ROOT/include -c $ROOT/charls_b200/csrc/host/$f.cpp -o $OUT/$kind/$f.o
  done
  runtime=$([ $kind = undefined ] && echo -lubsan || echo -lasan)
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/$kind/libcharls.so.3 $ROOT/charls_b200/build/jls_kernels.cu.o \
       $ROOT/charls_b200/build/engine.cu.o $OUT/$kind/*.o -Xlinker -soname,libcharls.so.3 \
       -Xlinker --version-script=$ROOT/charls_b200/csrc/exports.map -cudart static -Xcompiler -static-libgcc $runtime
  preload=""
  # ASan intercepts __cxa_throw and needs libstdc++ in front of the python executable, which does not link it
  [ $kind = address ] && preload="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libstdc++.so.6)"
  LD_PRELOAD="$preload" ASAN_OPTIONS=detect_leaks=0 CHARLS_B200_LIBRARY=$OUT/$kind/libcharls.so.3 \
      python -m pytest tests/test_abi.py tests/test_abi_differential.py -q -s 2>&1 | grep -E "runtime error|AddressSanitizer|passed|failed" || true
  LD_PRELOAD="$preload" ASAN_OPTIONS=detect_leaks=0 CHARLS_B200_LIBRARY=$OUT/$kind/libcharls.so.3 \
      python tools/parity_hunt.py abi 2>&1 | grep -E "runtime error|AddressSanitizer|hunt done" || true
done
