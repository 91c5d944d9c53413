#!/bin/bash
# usage: tools/ab_run.sh variant... ; prints value / encode / decode kernel ms for each variant (cfg2, 128 frames)
for v in "$@"; do
  CHARLS_B200_LIBRARY=$PWD/charls_b200/build/variants/$v/libcharls.so.3 python bench.py --steps 8 --no-cpu --no-e2e ${AB_ARGS} 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['value']), round(d['roofline_all']['encode']['ms_per_launch'],3), round(d['roofline_all']['decode']['ms_per_launch'],3))"
done
