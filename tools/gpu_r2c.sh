#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "replay or rng or launch_grouping or exact or smoke or singular" > gpurun_out/r2i_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2i_pytest.log
tail -6 gpurun_out/r2i_pytest.log
timeout 300 python bench.py --lattice 6 --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench_108_pi.json 2> gpurun_out/r2i_bench_108_pi.err
timeout 300 python bench.py --lattice 6 --steps 40 --warmup 3 --no-cpu-baseline --walkers-per-gpu 16384 --no-carlo > gpurun_out/r2i_bench_108_16k.json 2> gpurun_out/r2i_bench_108_16k.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_resident -s 1 -c 1 -o gpurun_out/prof_resident_r2i -f python tools/quick_bench.py --n 6 --walkers 4096 --sweeps 216 --therm 54 --no-prof > gpurun_out/ncu_resident_r2i.log 2>&1
for f in gpurun_out/r2i_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("metric","value","ms_per_step","kernel_ms")}, "e2e", (d.get("e2e") or {}).get("value"), "carlo", d.get("e2e_carlo"))
    r=d.get("roofline") or {}
    print("roofline", r.get("kernel","")[:30], r.get("frac"), r.get("hbm_GBs"), r.get("rank1_equivalent_GBs"), d.get("observables"))
except Exception as e:
    print("ERR", e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
done
