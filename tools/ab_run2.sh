#!/bin/bash
# usage: tools/ab_run2.sh variant... ; like ab_run.sh, plus the occupancy trace lines of the tile kernels (stderr)
for v in "$@"; do
  CHARLS_B200_TRACE_OCCUPANCY=1 CHARLS_B200_LIBRARY=$PWD/charls_b200/build/variants/$v/libcharls.so.3 python bench.py --steps 8 --no-cpu --no-e2e ${AB_ARGS} > /tmp/ab.out 2> /tmp/ab.err
  grep "resident blocks" /tmp/ab.err | sort -u | sed "s/^/$v /"
  tail -1 /tmp/ab.out | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', '${CHARLS_B200_CARVEOUT}', round(d['value']), round(d['roofline_all']['encode']['ms_per_launch'],3), round(d['roofline_all']['decode']['ms_per_launch'],3))"
done
