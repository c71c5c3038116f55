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
for cl in 4 5; do
timeout 300 python tools/quick_bench.py --n 18 --walkers 2048 --sweeps 486 --therm 486 --opt inverse_cluster=$cl > gpurun_out/r3p_q972_cl$cl.log 2>&1; echo "== 972 cl$cl"; show gpurun_out/r3p_q972_cl$cl.log
done
timeout 300 python tools/quick_bench.py --n 12 --B 0.02 --walkers 4096 --sweeps 216 --therm 216 > gpurun_out/r3p_qc128.log 2>&1; echo "== c128"; show gpurun_out/r3p_qc128.log
export KDSL_LIB=$PWD/kagomedsl.jl_b200/csrc/libkdsl_ticks.so
echo "== 972 cluster 4 phases"; timeout 300 python tools/cl_phases.py 18 1024 inverse_cluster=4 2>&1 | tail -4
