#!/bin/bash
mkdir -p gpurun_out
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
timeout 300 python tools/quick_bench.py --n 12 --walkers 4096 --sweeps 432 --therm 432 > gpurun_out/r3w_q432.log 2>&1; echo "== 432"; show gpurun_out/r3w_q432.log
timeout 300 python tools/quick_bench.py --n 18 --walkers 2048 --sweeps 486 --therm 486 > gpurun_out/r3w_q972.log 2>&1; echo "== 972"; show gpurun_out/r3w_q972.log
timeout 300 python tools/quick_bench.py --n 8 --walkers 4096 --sweeps 384 --therm 192 > gpurun_out/r3w_q192.log 2>&1; echo "== 192"; show gpurun_out/r3w_q192.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py -m gpu -x -q -k "replay or rng or full_batch or grouping" 2>&1 | tail -3
