#!/bin/bash
mkdir -p gpurun_out
export KDSL_DEBUG_OCC=1
run() {
  tag=$1; shift
  timeout 300 python tools/quick_bench.py --n 12 --B 0.02 --walkers 4096 --sweeps 216 --therm 216 "$@" > gpurun_out/r3h_qc128_$tag.log 2>&1
  echo "== c128 432 $tag"; grep -E "k_inverse_cl_c" gpurun_out/r3h_qc128_$tag.log | head -1
  python - gpurun_out/r3h_qc128_$tag.log <<'PY'
import sys,re,json
t=open(sys.argv[1]).read()
i=t.find('{'); j=t.rfind('}')
try:
    d=json.loads(t[i:j+1]); print(round(d["walker_sweeps_per_s"]/1e6,2), {k:(round(v["ms"],2),v["launches"]) for k,v in d["timers"].items()}, d["E_site"], d["n_singular"], round(d["flush_GBs"]))
except Exception as e: print("ERR",e, t[-800:])
PY
}
run nb8
run nb16x2 --opt inverse_tuning=2
run nb16x1 --opt inverse_tuning=3
run nb16x2_cl3 --opt inverse_tuning=2 --opt inverse_cluster=3
run nb16x2_cl2 --opt inverse_tuning=2 --opt inverse_cluster=2
run nb16x2_cl6 --opt inverse_tuning=2 --opt inverse_cluster=6
run nb16x2_rs2 --opt inverse_tuning=2 --opt inverse_row_slices=2
run nb16x2_rs4 --opt inverse_tuning=2 --opt inverse_row_slices=4
