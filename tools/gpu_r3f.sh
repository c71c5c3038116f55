#!/bin/bash
mkdir -p gpurun_out
run() {
  tag=$1; shift
  timeout 300 python tools/quick_bench.py --n 12 --B 0.02 --walkers 4096 --sweeps 216 --therm 216 "$@" > gpurun_out/r3f_qc128_$tag.log 2>&1
  echo "== c128 432 $tag"
  python - gpurun_out/r3f_qc128_$tag.log <<'PY'
import sys,re,json
t=open(sys.argv[1]).read()
i=t.find('{'); j=t.rfind('}')
try:
    d=json.loads(t[i:j+1]); print(round(d["walker_sweeps_per_s"]/1e6,2), {k:(round(v["ms"],2),v["launches"]) for k,v in d["timers"].items()}, d["E_site"], d["n_singular"])
except Exception as e: print("ERR",e, t[-800:])
PY
}
run g0
run g4 --opt gemm_variant=4
run g5 --opt gemm_variant=5
run g6 --opt gemm_variant=6
