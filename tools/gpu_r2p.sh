#!/bin/bash
for o in "flush_variant=4" "flush_variant=4 --opt flush_dbg=1" "flush_variant=0 --opt flush_dbg=1"; do
echo "== $o"
python tools/quick_bench.py --n 12 --walkers 4096 --sweeps 432 --therm 432 --opt $o 2>&1 | grep -E "flush_GBs|walker_sweeps_per_s"
done
