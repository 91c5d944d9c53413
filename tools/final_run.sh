#!/bin/bash
# Round-end evidence on one B200 (run under gpurun): GPU tests, smoke, bench lines for every workload, the reference arm,
# the ncu launch list of a bench step and full captures of the two coder kernels (cfg2 and cfg4).  Outputs land in gpurun_out/.
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 700 gpurun_out/bench_final.json
python bench.py --workload cfg3 --frames 64 --steps 5 --no-cpu --also none > gpurun_out/bench_final_cfg3.json 2>/dev/null
python bench.py --workload cfg4 --frames 64 --steps 5 --no-cpu --also none > gpurun_out/bench_final_cfg4.json 2>/dev/null
python bench.py --frames 1 --steps 20 --no-cpu --no-e2e --also none > gpurun_out/bench_final_single.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --also none > gpurun_out/b_under_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_(encode|decode)_tiled" -s 6 -c 2 -o gpurun_out/prof_final -f python bench.py --frames 128 --steps 1 --warmup 3 --no-e2e --no-cpu --also none > gpurun_out/b_under_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_(encode|decode)_tiled" -s 6 -c 2 -o gpurun_out/prof_final_cfg4 -f python bench.py --workload cfg4 --frames 64 --steps 1 --warmup 3 --no-e2e --no-cpu --also none > gpurun_out/b_under_ncu_cfg4.log 2>&1
ls -la gpurun_out/prof_final.ncu-rep gpurun_out/prof_final_cfg4.ncu-rep gpurun_out/launches_final.csv
