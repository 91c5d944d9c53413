AB_WORKLOADS="cfg2:128" timeout 900 tools/r2_ab.sh 10 cur cursub bulk cur cursub 2>&1 | tail -6
python tools/sass_grep.py charls_b200/build/variants/bulk/libcharls.so.3 2>/dev/null | tail -3
