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
for mb in 0 48 80 112; do
timeout 300 python tools/quick_bench.py --n 12 --walkers 4096 --sweeps 432 --therm 432 --opt l2_persist_mb=$mb > gpurun_out/r3s_q432_p$mb.log 2>&1; echo "== 432 persist $mb MB"; grep "l2 persist" gpurun_out/r3s_q432_p$mb.log | head -1; show gpurun_out/r3s_q432_p$mb.log
done
for mb in 0 80; do
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_reeval_fused -s 2 -c 1 python tools/prof_refresh.py 12 1024 3 l2_persist_mb=$mb 2>&1 | grep -E "dram__bytes|gpu__time" | sed "s/^/persist $mb: /"
done
