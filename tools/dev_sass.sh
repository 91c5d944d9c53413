#!/bin/bash
# Development loop without a GPU: compiles the one- and three-component kernels only (JLS_DEV_SUBSET, ~half the time of the
# full build) into /tmp/jls_dev.o and prints the straight-path instruction counts of the pixel loops (tools/sass_path.py).
# usage: tools/dev_sass.sh [-DJLS_...=v ...]
set -e
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden -DCHARLS_B200_BUILD \
  -DJLS_DEV_SUBSET "$@" -Iinclude -Icharls_b200/csrc -x cu -c charls_b200/csrc/jls_kernels.cu -o /tmp/jls_dev.o
for k in k_encode_tiledILi1ELb1EhLi8 k_decode_tiledILi1ELb1EhLi8 k_encode_tiledILi1ELb0EtLi0 k_decode_tiledILi1ELb0EtLi0; do
  python tools/sass_path.py /tmp/jls_dev.o $k --quiet
done
for k in k_encode_tiledILi3ELb1EtLi16 k_decode_tiledILi3ELb1EtLi16; do
  python tools/sass_path.py /tmp/jls_dev.o $k --quiet --passes 3
done
cuobjdump --dump-resource-usage /tmp/jls_dev.o | grep -A1 -E "tiledILi[13]ELb[01]E[ht]Li(8|16|0)E" | grep -E "Function|REG" | paste - - \
  | sed -E 's/.*(k_[a-z]+_tiledILi.ELb.E.Li[0-9]+).*REG:([0-9]+) STACK:([0-9]+) SHARED:([0-9]+).*/\1 REG \2 STACK \3 SHARED \4/' | grep -E "Lb1EhLi8|Lb0EtLi0|ILi3ELb1EtLi16"
