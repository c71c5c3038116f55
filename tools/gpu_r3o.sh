#!/bin/bash
mkdir -p gpurun_out
export KDSL_DEBUG_OCC=1
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py -m gpu -x -q -k "complex or c128 or cluster or 972 or variants or invariants" > gpurun_out/r3o_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3o_pytest.log
tail -6 gpurun_out/r3o_pytest.log
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
for cl in 4 5 6; do
timeout 300 python tools/quick_bench.py --n 18 --walkers 2048 --sweeps 486 --therm 486 --opt inverse_cluster=$cl > gpurun_out/r3o_q972_cl$cl.log 2>&1; echo "== 972 cl$cl"; show gpurun_out/r3o_q972_cl$cl.log
done
for cl in 4 5; do
timeout 300 python tools/quick_bench.py --n 12 --B 0.02 --walkers 4096 --sweeps 216 --therm 216 --opt inverse_cluster=$cl > gpurun_out/r3o_qc128_cl$cl.log 2>&1; echo "== c128 cl$cl"; grep k_inverse_cl_c gpurun_out/r3o_qc128_cl$cl.log | head -1; show gpurun_out/r3o_qc128_cl$cl.log
done
