#!/bin/bash
mkdir -p gpurun_out
for sel in "refresh_matches_oracle" "imbalanced" "complex_refresh"; do
  f=gpurun_out/r4c_race_$(echo $sel | tr ' ' '_').log
  timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$sel" > $f 2>&1
  echo "== $sel: $(grep -c 'hazard' $f) hazard lines"; grep -E "RACECHECK SUMMARY|passed|failed" $f | tail -2
  grep -E "Race reported|hazard detected|at .*\(|in k_|void k_" $f | sort | uniq -c | sort -rn | head -12
done
