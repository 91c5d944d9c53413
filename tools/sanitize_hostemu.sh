#!/bin/bash
# The kernels' codec code (jls_codec.cuh, jls_fast.cuh, jls_interval.cuh) compiled for the host with UBSan, then ASan, and run
# through tests/test_hostemu.py (valid and damaged lines).  Restores the normal library afterwards.
# End of round 1: one finding (signed overflow of A on damaged input after the sanity mark had tripped), fixed; ASan clean.
set -e
cd "$(dirname "$0")/.."
build() { g++ -O1 -g -std=c++17 -fPIC -shared "$@" -Icharls_b200/csrc -o tests/hostemu/libhostemu.so tests/hostemu/hostemu.cpp; }
build -fsanitize=undefined
python -m pytest tests/test_hostemu.py -q -s 2>&1 | grep -E "runtime error|passed|failed" || true
build -fsanitize=address
# libstdc++ behind libasan: the damaged walk runs the reference beside the host build, and the reference throws
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libstdc++.so.6)" ASAN_OPTIONS=detect_leaks=0 python -m pytest tests/test_hostemu.py -q -s 2>&1 | grep -E "AddressSanitizer|passed|failed" || true
build -O2
# the oracle itself (the checker must not lean on undefined behaviour either): two overflow sites on damaged input were
# found and made to wrap explicitly at the end of round 1
preload="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libstdc++.so.6)"
gcc -O1 -g -fPIC -std=c11 -fsanitize=undefined -shared -o oracle/libjls_oracle.so oracle/jls_oracle.c -lubsan
python -m pytest tests/test_oracle.py tests/test_hostemu.py -q -s 2>&1 | grep -E "runtime error|passed|failed" || true
gcc -O1 -g -fPIC -std=c11 -fsanitize=address -shared -o oracle/libjls_oracle.so oracle/jls_oracle.c
LD_PRELOAD="$preload" ASAN_OPTIONS=detect_leaks=0 python -m pytest tests/test_oracle.py tests/test_hostemu.py -q -s 2>&1 | grep -E "AddressSanitizer|passed|failed" || true
make -s -C oracle clean && make -s -C oracle libjls_oracle.so
