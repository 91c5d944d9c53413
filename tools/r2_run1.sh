python -m pytest tests -x -q -m gpu 2>&1 | tail -5
tools/r2_ab.sh 1 r1 cur cache1 fma0 steady0 bias0 2>&1 | tail -20
