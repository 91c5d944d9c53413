timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 tools/r2_ab.sh 9 r2a topup4 cur 2>&1 | tail -9
