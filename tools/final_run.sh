python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/bench7.json 2> gpurun_out/bench7.err; tail -c 600 gpurun_out/bench7.json
python bench.py --workload cfg3 --frames 64 --steps 5 --no-cpu > gpurun_out/bench7_cfg3.json 2>/dev/null
python bench.py --workload cfg4 --frames 64 --steps 5 --no-cpu > gpurun_out/bench7_cfg4.json 2>/dev/null
python bench.py --frames 1 --steps 20 --no-cpu --no-e2e > gpurun_out/bench7_single.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench7_reference.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1g.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/b_under_ncu7.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_(encode|decode)_tiled" -s 6 -c 2 -o gpurun_out/prof_r1g -f python bench.py --frames 128 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_under_ncu8.log 2>&1
ls -la gpurun_out/prof_r1g.ncu-rep gpurun_out/launches_r1g.csv
