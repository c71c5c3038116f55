#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "trajectory or rng or full_batch or zero_flux or energy_12 or reweighted or observables or smoke or frozen or resume" 2>&1 | tail -4
for i in 1 2; do python bench.py --lattice 6 --steps 40 --warmup 3 --no-cpu-baseline --no-carlo --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['metric'], round(d['value']/1e6,2), d['ms_per_step'], d['observables'])"; done
python bench.py --lattice 6 --flux zero --steps 40 --warmup 3 --no-cpu-baseline --no-carlo 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['metric'], round(d['value']/1e6,2), d['ms_per_step'], 'e2e', round(d['e2e']['value']/1e6,2))"
