timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 tools/r2_ab.sh 7 r2a wide cur 2>&1 | tail -9
CHARLS_B200_TRACE_OCCUPANCY=1 timeout 100 python bench.py --steps 2 --no-cpu --no-e2e --also none --workload cfg4 --frames 16 2>&1 | grep "resident" | sort -u
