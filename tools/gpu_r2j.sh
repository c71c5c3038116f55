#!/bin/bash
mkdir -p gpurun_out
T=${1:-r2j}
timeout 900 python -m pytest tests -m gpu -x -q -k "replay_432 or replay_972 or launch_grouping or trajectory or pending or rng" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
tail -6 gpurun_out/${T}_pytest.log
for fv in 0 2; do
  timeout 300 python tools/quick_bench.py --n 12 --walkers 4096 --sweeps 432 --therm 432 --opt flush_variant=$fv > gpurun_out/${T}_qb432_fv$fv.log 2>&1
  timeout 600 python tools/quick_bench.py --n 18 --walkers 2048 --sweeps 486 --therm 486 --opt flush_variant=$fv > gpurun_out/${T}_qb972_fv$fv.log 2>&1
done
for f in gpurun_out/${T}_qb*.log; do echo "== $f"; grep -E "walker_sweeps_per_s|flush_GBs|\"update\"|\"propose\"" -A2 $f | grep -E "walker_sweeps|flush_GBs|ms|flushes" | tr -d '\n'; echo; done
