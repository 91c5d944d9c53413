timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 100 python tools/prof_general.py 2>&1 | tail -2
timeout 600 ncu --set full --import-source on --clock-control none -k regex:general -s 2 -c 2 -o /tmp/prof_general -f python tools/prof_general.py > gpurun_out/prof_general.log 2>&1
ncu -i /tmp/prof_general.ncu-rep --page raw --csv > gpurun_out/prof_general_raw.csv 2>/dev/null
ncu -i /tmp/prof_general.ncu-rep --page source --csv --kernel-name regex:k_decode_general > gpurun_out/prof_general_decode_source.csv 2>/dev/null
ls -la gpurun_out/prof_general*
