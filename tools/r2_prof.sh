#!/bin/bash
# usage (under gpurun): tools/r2_prof.sh <tag> [workload frames]
# Full ncu capture of the two coder kernels of the in-tree library.  The report stays on the box (two of them exceed the
# 64 MiB that travel back); the raw metric page and the per-instruction source page come back as CSV.
tag=$1; wl=${2:-cfg2}; fr=${3:-128}
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_(encode|decode)_tiled" -s 6 -c 2 -o /tmp/prof_$tag -f \
  python bench.py --workload $wl --frames $fr --steps 1 --warmup 3 --no-e2e --no-cpu --also none > gpurun_out/prof_$tag.log 2>&1
ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2>/dev/null
for k in encode decode; do
  ncu -i /tmp/prof_$tag.ncu-rep --page source --csv --kernel-name regex:k_${k}_tiled > gpurun_out/prof_${tag}_${k}_source.csv 2>/dev/null
done
ls -la gpurun_out/prof_${tag}_*
