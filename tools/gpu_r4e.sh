#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do python tools/quick_bench.py --n 6 --walkers 4096 --sweeps 432 --therm 108 2>&1 | grep -E "walker_sweeps_per_s"; done
f=gpurun_out/r4e_race_traj.log
timeout 2400 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "trajectory_bit_exact" > $f 2>&1
grep -E "RACECHECK SUMMARY|passed|failed" $f | tail -2
grep -E "Race reported" $f | sed 's/0x[0-9a-f]*/ADDR/g' | cut -c1-260 | sort | uniq -c | sort -rn | head -10
