#!/bin/bash
timeout 600 python -m pytest tests -m gpu -x -q -k "replay_432 or replay_972 or launch_grouping or trajectory or pending" 2>&1 | tail -3
for o in "flush_variant=2" "flush_variant=3" "flush_variant=4" "flush_variant=5" "flush_variant=4 --opt flush_dbg=1"; do
echo "== 432 $o"
python tools/quick_bench.py --n 12 --walkers 4096 --sweeps 432 --therm 432 --opt $o 2>&1 | grep -E "flush_GBs|walker_sweeps_per_s" | tr -d '\n'; echo
done
for o in "flush_variant=2" "flush_variant=3" "flush_variant=4"; do
echo "== 972 $o"
python tools/quick_bench.py --n 18 --walkers 2048 --sweeps 486 --therm 486 --opt $o 2>&1 | grep -E "flush_GBs|walker_sweeps_per_s" | tr -d '\n'; echo
done
