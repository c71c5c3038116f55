#!/bin/bash
# round 2, first GPU call: gpu tests, headline bench, other-config baselines, launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench_432.json 2> gpurun_out/r2a_bench_432.err; tail -c 600 gpurun_out/r2a_bench_432.err
python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
python bench.py --lattice 18 --steps 3 --warmup 3 --no-cpu-baseline --no-carlo > gpurun_out/r2a_bench_972.json 2> gpurun_out/r2a_bench_972.err
python bench.py --lattice 6 --steps 40 --warmup 3 --no-cpu-baseline --no-carlo > gpurun_out/r2a_bench_108.json 2> gpurun_out/r2a_bench_108.err
python bench.py --lattice 6 --flux zero --steps 40 --warmup 3 --no-cpu-baseline --no-carlo > gpurun_out/r2a_bench_108_zero.json 2> gpurun_out/r2a_bench_108_zero.err
python bench.py --lattice 12 --B 0.02 --steps 3 --warmup 3 --no-cpu-baseline --no-carlo > gpurun_out/r2a_bench_432_c128.json 2> gpurun_out/r2a_bench_432_c128.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2a_launches.log 2>&1
for f in gpurun_out/r2a_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("metric","value","ms_per_step","kernel_ms")}, "e2e", (d.get("e2e") or {}).get("value"), "carlo", d.get("e2e_carlo"), "reduce", d.get("reduce"))
    r=d.get("roofline") or {}
    print("roofline", r.get("kernel","")[:30], r.get("frac"), r.get("frac_contract"), "upd", (d.get("roofline_w_update") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value_f64"), "dmma", d.get("dmma_probe"))
except Exception as e:
    print("ERR", e)
PY
done
