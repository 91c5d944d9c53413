timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
AB_WORKLOADS="cfg2:128 cfg4:64" timeout 600 tools/r2_ab.sh 6 r2a cur nodiscard 2>&1 | tail -8
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:^k_ -c 60 --csv --log-file gpurun_out/launches_r2b.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --also none > gpurun_out/launches_r2b.log 2>&1
tail -30 gpurun_out/launches_r2b.csv | cut -c1-200
