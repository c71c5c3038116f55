#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py -m gpu -x -q -k "complex or c128" > gpurun_out/r3g_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3g_pytest.log
tail -15 gpurun_out/r3g_pytest.log
run() {
  tag=$1; shift
  timeout 300 python tools/quick_bench.py --n 12 --B 0.02 --walkers 4096 --sweeps 216 --therm 216 "$@" > gpurun_out/r3g_qc128_$tag.log 2>&1
  echo "== c128 432 $tag"
  python - gpurun_out/r3g_qc128_$tag.log <<'PY'
import sys,re,json
t=open(sys.argv[1]).read()
i=t.find('{'); j=t.rfind('}')
try:
    d=json.loads(t[i:j+1]); print(round(d["walker_sweeps_per_s"]/1e6,2), {k:(round(v["ms"],2),v["launches"]) for k,v in d["timers"].items()}, d["E_site"], d["n_singular"], round(d["flush_GBs"]))
except Exception as e: print("ERR",e, t[-800:])
PY
}
run f0
run f4 --opt flush_variant=4
