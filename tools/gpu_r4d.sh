#!/bin/bash
mkdir -p gpurun_out
f=gpurun_out/r4d_race_traj.log
timeout 2400 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "trajectory_bit_exact and 2-2" > $f 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" $f | tail -2
grep -E "Race reported" $f | sed 's/0x[0-9a-f]*/ADDR/g' | sort | uniq -c | sort -rn | head -20
grep -E "^=========     (Write|Read) Thread|in .*k_[a-z_]*" $f | sed 's/(.*//' | sort | uniq -c | sort -rn | head -12
head -c 3000 $f | tail -c 2200
