#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3y_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3y_pytest.log
tail -5 gpurun_out/r3y_pytest.log
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
run() { tag=$1; n=$2; shift; shift; timeout 300 python tools/quick_bench.py --n $n --walkers 4096 "$@" > gpurun_out/r3y_$tag.log 2>&1; echo "== $tag"; show gpurun_out/r3y_$tag.log; }
run 192_small 8 --sweeps 384 --therm 192
run 192_big 8 --sweeps 384 --therm 192 --opt inverse_tuning=7
run 108_lockstep_small 6 --sweeps 432 --therm 108 --opt update_variant=2
run 108_lockstep_big 6 --sweeps 432 --therm 108 --opt update_variant=2 --opt inverse_tuning=7
run 432 12 --sweeps 432 --therm 432
