#!/bin/bash
# usage (under gpurun): tools/sanitize_gpu.sh [tool ...]     default: memcheck racecheck initcheck synccheck
# Runs tools/sanitize_workload.py under compute-sanitizer, one tool at a time; summaries land in gpurun_out/sanitize_<tool>.txt.
tools=("$@"); [ ${#tools[@]} -eq 0 ] && tools=(memcheck racecheck initcheck synccheck)
for t in "${tools[@]}"; do
  extra=""
  [ "$t" = memcheck ] && extra="--leak-check no"
  timeout 900 compute-sanitizer --tool $t $extra --error-exitcode 86 --print-limit 30 python tools/sanitize_workload.py > gpurun_out/sanitize_$t.txt 2>&1
  echo "$t: exit $? | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize workload' gpurun_out/sanitize_$t.txt | tr '\n' ' ')"
done
