#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "launch_grouping" 2>&1 | tail -3
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
for fv in 0 5 6; do
timeout 300 python tools/quick_bench.py --n 12 --walkers 4096 --sweeps 432 --therm 432 --opt flush_variant=$fv > gpurun_out/r4j_q432_$fv.log 2>&1; echo "== 432 fv $fv"; show gpurun_out/r4j_q432_$fv.log
timeout 300 python tools/quick_bench.py --n 18 --walkers 2048 --sweeps 486 --therm 486 --opt flush_variant=$fv > gpurun_out/r4j_q972_$fv.log 2>&1; echo "== 972 fv $fv"; show gpurun_out/r4j_q972_$fv.log
done
