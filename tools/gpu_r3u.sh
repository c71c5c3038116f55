#!/bin/bash
mkdir -p gpurun_out
export KDSL_DEBUG_OCC=1
show() {
  python - $1 <<'PY'
import sys,re,json
t=open(sys.argv[1]).read()
i=t.find('{'); j=t.rfind('}')
try:
    d=json.loads(t[i:j+1]); print(round(d["walker_sweeps_per_s"]/1e6,2), {k:(round(v["ms"],2),v["launches"]) for k,v in d["timers"].items()}, d["E_site"], d["n_singular"], round(d["flush_GBs"]))
except Exception as e: print("ERR",e, t[-1200:])
PY
}
run() { tag=$1; shift; timeout 600 python tools/quick_bench.py --n 18 --B 0.01 --walkers 1024 --sweeps 486 --therm 486 "$@" > gpurun_out/r3u_qc972_$tag.log 2>&1; echo "== c128 972 $tag"; grep "k_inverse_cl_c" gpurun_out/r3u_qc972_$tag.log | head -1; show gpurun_out/r3u_qc972_$tag.log; }
run def
run f4 --opt flush_variant=4
run g4 --opt gemm_variant=4
run v7 --opt inverse_variant=7
