#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/config_sweep.py 2>&1 | tail -14
python bench.py --lattice 6 --flux zero --steps 40 --warmup 3 --no-cpu-baseline --no-carlo > gpurun_out/r3r_bench_108_zero.json 2> gpurun_out/r3r_bench_108_zero.err
python bench.py --lattice 8 --B 0.02 --walkers-per-gpu 16384 --steps 5 --warmup 3 --no-cpu-baseline --no-carlo > gpurun_out/r3r_bench_192_c128.json 2> gpurun_out/r3r_bench_192_c128.err
python bench.py --lattice 12 --B 0.02 --walkers-per-gpu 16384 --steps 3 --warmup 3 --no-cpu-baseline --no-carlo > gpurun_out/r3r_bench_432_c128_16k.json 2> gpurun_out/r3r_bench_432_c128_16k.err
for f in gpurun_out/r3r_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d["metric"], round(d["value"]/1e6,2), d["ms_per_step"], d["observables"])
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
