#!/bin/bash
# round 2: resident kernel validation
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_pytest.log
tail -15 gpurun_out/r2b_pytest.log
for fl in pi zero; do
timeout 300 python bench.py --lattice 6 --flux $fl --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_108_$fl.json 2> gpurun_out/r2b_bench_108_$fl.err
done
timeout 300 python bench.py --lattice 6 --steps 40 --warmup 3 --no-cpu-baseline --walkers-per-gpu 16384 --no-carlo > gpurun_out/r2b_bench_108_16k.json 2> gpurun_out/r2b_bench_108_16k.err
timeout 300 python bench.py --lattice 6 --steps 40 --warmup 3 --no-cpu-baseline --no-carlo --options update_variant=2 > gpurun_out/r2b_bench_108_wb.json 2> gpurun_out/r2b_bench_108_wb.err
for f in gpurun_out/r2b_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
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
