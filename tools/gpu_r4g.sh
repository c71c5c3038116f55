#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r4g_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r4g_pytest.log
tail -4 gpurun_out/r4g_pytest.log
python bench.py --lattice 12 --B 0.02 --steps 5 --warmup 3 --no-cpu-baseline --no-carlo > gpurun_out/r4g_bench_432_c128.json 2> gpurun_out/r4g_bench_432_c128.err
python bench.py --lattice 8 --steps 20 --warmup 3 --no-cpu-baseline --no-carlo > gpurun_out/r4g_bench_192.json 2> gpurun_out/r4g_bench_192.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r4g_bench_432.json 2> gpurun_out/r4g_bench_432.err
for f in gpurun_out/r4g_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=d.get("roofline") or {}
    print(sys.argv[1], d["metric"], round(d["value"]/1e6,2), round(d["ms_per_step"],3), "e2e", round((d.get("e2e") or {}).get("value",0)/1e6,2), r.get("kernel","")[:28], round(r.get("frac") or 0,3), "upd", round((d.get("roofline_w_update") or {}).get("frac") or 0,3), "gemm", (d.get("roofline_refresh_gemm") or {}).get("frac"))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
