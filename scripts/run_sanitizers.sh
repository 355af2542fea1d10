#!/bin/bash
# compute-sanitizer racecheck + synccheck (+ memcheck) over small instances of every hot kernel; summaries under gpurun_out/sanitizer/.
# Usage (on the GPU box): bash scripts/run_sanitizers.sh [cases...]
out=gpurun_out/sanitizer
mkdir -p $out
cases=${@:-"bamp_c1 bamp_c2 bamp_c2_na4 vamp_c2 vamp_c3 vamp_from_h scamp_tc scamp_simt bamp_generic"}
for c in $cases; do
  for tool in racecheck synccheck memcheck; do
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_case.py --case $c > $out/${c}.${tool}.log 2>&1
    echo "$c $tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|hazard' $out/${c}.${tool}.log | tail -2 | tr '\n' ' ')"
  done
done | tee $out/summary.txt
