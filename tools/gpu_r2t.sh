#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2t_pytest.log
tail -8 gpurun_out/r2t_pytest.log
for o in "" "--opt update_variant=0"; do
echo "== c128 432 $o"
timeout 300 python tools/quick_bench.py --n 12 --walkers 4096 --sweeps 432 --therm 432 --B 0.02 $o 2>&1 | grep -E "walker_sweeps_per_s|flush_GBs|update_GBs|\"ms\"|E_site" | tr -d '\n'; echo
done
timeout 300 python bench.py --lattice 12 --B 0.02 --steps 5 --warmup 3 --no-cpu-baseline --no-carlo > gpurun_out/r2t_bench_432_c128.json 2> gpurun_out/r2t_bench_432_c128.err; tail -c 300 gpurun_out/r2t_bench_432_c128.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2t_bench_432_c128.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("metric","value","ms_per_step","kernel_ms")}, (d.get("e2e") or {}).get("value"), d.get("observables"))
    print(d["roofline"]["kernel"][:40], d["roofline"]["frac"], d["roofline_w_update"]["frac"])
except Exception as e: print("ERR", e)
PY
