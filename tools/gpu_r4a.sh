#!/bin/bash
mkdir -p gpurun_out
show() {
  python - $1 <<'PY'
import sys,re,json
t=open(sys.argv[1]).read()
i=t.find('{'); j=t.rfind('}')
try:
    d=json.loads(t[i:j+1]); print(round(d["walker_sweeps_per_s"]/1e6,2), {k:(round(v["ms"],2),v["launches"]) for k,v in d["timers"].items()}, d["E_site"], round(d["flush_GBs"]))
except Exception as e: print("ERR",e, t[-800:])
PY
}
for dbg in 0 2 4; do
timeout 300 python tools/quick_bench.py --n 12 --walkers 4096 --sweeps 432 --therm 432 --opt flush_dbg=$dbg > gpurun_out/r4a_q432_$dbg.log 2>&1; echo "== 432 RT dbg $dbg"; show gpurun_out/r4a_q432_$dbg.log
timeout 300 python tools/quick_bench.py --n 18 --walkers 2048 --sweeps 486 --therm 486 --opt flush_dbg=$dbg > gpurun_out/r4a_q972_$dbg.log 2>&1; echo "== 972 RT dbg $dbg"; show gpurun_out/r4a_q972_$dbg.log
done
