#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py -m gpu -x -q -k "variants or cluster or 972 or invariants" > gpurun_out/r3j_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3j_pytest.log
tail -8 gpurun_out/r3j_pytest.log
run() {
  tag=$1; shift
  timeout 300 python tools/quick_bench.py --n 18 --walkers 2048 --sweeps 486 --therm 486 "$@" > gpurun_out/r3j_q972_$tag.log 2>&1
  echo "== 972 $tag"
  python - gpurun_out/r3j_q972_$tag.log <<'PY'
import sys,re,json
t=open(sys.argv[1]).read()
i=t.find('{'); j=t.rfind('}')
try:
    d=json.loads(t[i:j+1]); print(round(d["walker_sweeps_per_s"]/1e6,2), {k:(round(v["ms"],2),v["launches"]) for k,v in d["timers"].items()}, d["E_site"], d["n_singular"], round(d["flush_GBs"]))
except Exception as e: print("ERR",e, t[-800:])
PY
}
run g0
run g1 --opt gemm_variant=1
run g5 --opt gemm_variant=5
run g6 --opt gemm_variant=6
