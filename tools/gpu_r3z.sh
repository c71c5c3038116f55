#!/bin/bash
# round 2, session 3: full gpu suite, bench lines for every BASELINE config with the session's kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3z_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3z_pytest.log
tail -5 gpurun_out/r3z_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r3z_bench_432.json 2> gpurun_out/r3z_bench_432.err; tail -c 400 gpurun_out/r3z_bench_432.err
python bench.py --lattice 18 --steps 3 --warmup 3 --no-cpu-baseline --no-carlo > gpurun_out/r3z_bench_972.json 2> gpurun_out/r3z_bench_972.err
python bench.py --lattice 6 --steps 40 --warmup 3 --no-cpu-baseline --no-carlo > gpurun_out/r3z_bench_108.json 2> gpurun_out/r3z_bench_108.err
python bench.py --lattice 12 --B 0.02 --steps 5 --warmup 3 --no-cpu-baseline --no-carlo > gpurun_out/r3z_bench_432_c128.json 2> gpurun_out/r3z_bench_432_c128.err; tail -c 300 gpurun_out/r3z_bench_432_c128.err
for f in gpurun_out/r3z_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("metric","value","ms_per_step","kernel_ms")}, "e2e", (d.get("e2e") or {}).get("value"))
    r=d.get("roofline") or {}
    print("roofline", r.get("kernel","")[:30], r.get("frac"), r.get("frac_contract"), "upd", (d.get("roofline_w_update") or {}).get("frac"), "inv", (d.get("roofline_refresh") or {}).get("frac"), "gemm", (d.get("roofline_refresh_gemm") or {}).get("frac"))
except Exception as e:
    print("ERR", e)
PY
done
