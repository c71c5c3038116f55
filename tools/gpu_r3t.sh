#!/bin/bash
mkdir -p gpurun_out
export KDSL_DEBUG_OCC=1
show() {
  python - $1 <<'PY'
import sys,re,json
t=open(sys.argv[1]).read()
i=t.find('{'); j=t.rfind('}')
try:
    d=json.loads(t[i:j+1]); print(round(d["walker_sweeps_per_s"]/1e6,2), {k:(round(v["ms"],2),v["launches"]) for k,v in d["timers"].items()}, d["E_site"], d["n_singular"])
except Exception as e: print("ERR",e, t[-800:])
PY
}
run() { tag=$1; shift; timeout 300 python tools/quick_bench.py --n 8 --walkers 4096 --sweeps 384 --therm 192 "$@" > gpurun_out/r3t_q192_$tag.log 2>&1; echo "== 192 $tag"; grep k_reeval_cl gpurun_out/r3t_q192_$tag.log | head -1; show gpurun_out/r3t_q192_$tag.log; }
run fused
run cl4 --opt inverse_variant=8
run cl3 --opt inverse_variant=8 --opt reeval_cluster=3
run cl2 --opt inverse_variant=8 --opt reeval_cluster=2
run v5 --opt inverse_variant=5
run v7 --opt inverse_variant=7
