#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q -k "complex or c128" 2>&1 | tail -3
for fv in 0 7; do python tools/quick_bench.py --n 12 --B 0.02 --walkers 4096 --sweeps 216 --therm 216 --opt flush_variant=$fv 2>&1 | grep -E "walker_sweeps_per_s|E_site" | tr -d "\n"; python tools/quick_bench.py --n 12 --B 0.02 --walkers 4096 --sweeps 216 --therm 216 --opt flush_variant=$fv 2>&1 | grep -A1 '"update"' | tr -d "\n"; echo " fv=$fv"; done
