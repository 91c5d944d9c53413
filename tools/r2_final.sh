#!/bin/bash
TAG=${1:-r2k}
# Round-2 evidence on one B200 (run under gpurun): GPU tests, smoke, bench lines, the ncu launch list of a bench step and full
# captures of the two coder kernels (cfg2 and cfg4).  Reports stay on the box; CSV pages and bench lines land in gpurun_out/.
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; tail -c 400 gpurun_out/${TAG}_bench_default.json; echo
python bench.py --frames 1 --steps 20 --no-cpu --no-e2e --also none > gpurun_out/${TAG}_bench_single.json 2>/dev/null
python bench.py --content noise --steps 5 --no-cpu --no-e2e --also none > gpurun_out/${TAG}_bench_noise.json 2>/dev/null
python bench.py --content flat --steps 5 --no-cpu --no-e2e --also none > gpurun_out/${TAG}_bench_flat.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 200 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --also none > gpurun_out/${TAG}_under_ncu_launches.log 2>&1
bash tools/r2_prof.sh ${TAG} cfg2 128
bash tools/r2_prof.sh ${TAG}_cfg4 cfg4 128
for t in 24 32; do python bench.py --steps 5 --no-cpu --also none --e2e-one-part-threads $t > gpurun_out/${TAG}_bench_e2e_t$t.json 2>/dev/null; done
for f in single noise flat reference e2e_t24 e2e_t32; do python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_$f.json')); print('$f', round(d['value']), d.get('ms_per_step'), (d.get('e2e') or {}).get('value'))"; done
