#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q -k "measure or replay or full_batch or rng or energy_12 or observables" 2>&1 | tail -3
python tools/quick_bench.py --n 12 --walkers 4096 --sweeps 432 --therm 432 2>&1 | grep -E "walker_sweeps_per_s|E_site" ; python tools/quick_bench.py --n 12 --walkers 4096 --sweeps 432 --therm 432 2>&1 | grep -A1 '"measure"' | tr -d "\n"; echo
python tools/quick_bench.py --n 18 --walkers 2048 --sweeps 486 --therm 486 2>&1 | grep -A1 '"measure"' | tr -d "\n"; echo
