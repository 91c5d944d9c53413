#!/bin/bash
# usage (under gpurun): tools/r2_hints.sh   -- A/B of the L2 eviction hints (JLS_L2_HINTS, variants h0..h3 from tools/ab_build.py):
# bench numbers and the encoder's DRAM bytes per launch for each variant.
AB_WORKLOADS="cfg2:128 cfg4:128" bash tools/r2_ab.sh hints h0 h1 h2 h3
for v in h0 h1 h2 h3; do
  for wl in cfg2 cfg4; do
    CHARLS_B200_LIBRARY=$PWD/charls_b200/build/variants/$v/libcharls.so.3 timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
      --clock-control none -k regex:"k_(encode|decode)_tiled" -s 6 -c 2 --csv python bench.py --workload $wl --frames 128 --steps 1 --warmup 3 --no-e2e --no-cpu --also none 2>/dev/null \
      | grep -E "dram__bytes|gpu__time" | awk -F'","' -v v=$v -v wl=$wl '{print wl, v, $5, $(NF-2), $(NF-1), $NF}' | tr -d '"' >> gpurun_out/ab_hints_dram.txt
  done
done
cat gpurun_out/ab_hints_dram.txt
