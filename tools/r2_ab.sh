#!/bin/bash
# usage (under gpurun): tools/r2_ab.sh <tag> variant...   ("cur" = the in-tree library)
# A/B of library variants (tools/ab_build.py) on one box: value / encode / decode kernel ms for cfg2, cfg4 and cfg3.
# AB_WORKLOADS="cfg2:128 cfg4:64" restricts the workloads.
tag=$1; shift
variants=("$@")
out=gpurun_out/ab_$tag.txt
: > $out
for w in ${AB_WORKLOADS:-cfg2:128 cfg4:64 cfg3:64}; do
  wl=${w%%:*}; fr=${w##*:}
  for v in "${variants[@]}"; do
    lib=$PWD/charls_b200/build/variants/$v/libcharls.so.3
    [ "$v" = cur ] && lib=$PWD/charls_b200/lib/libcharls.so.3
    CHARLS_B200_LIBRARY=$lib python bench.py --workload $wl --frames $fr --steps 8 --no-cpu --no-e2e --also none 2>/tmp/ab.err | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$wl', '$v', round(d['value']), round(d['roofline_all']['encode']['ms_per_launch'],3), round(d['roofline_all']['decode']['ms_per_launch'],3), round(d['ms_per_step'],3))
except Exception as e:
    print('$wl', '$v', 'FAILED', e); print(open('/tmp/ab.err').read()[-600:])
" >> $out
  done
done
cat $out
