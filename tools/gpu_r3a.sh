#!/bin/bash
# round-2 session 2: cluster inverse bring-up (parity first, then 972-site timing with and without clusters)
mkdir -p gpurun_out
export KDSL_DEBUG_OCC=1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cluster or imbalanced or complex_refresh" > gpurun_out/r3a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3a_pytest.log
tail -15 gpurun_out/r3a_pytest.log
for cl in 4 0 5 3; do
  timeout 300 python tools/quick_bench.py --n 18 --walkers 2048 --sweeps 486 --therm 486 --opt inverse_cluster=$cl > gpurun_out/r3a_q972_cl$cl.log 2>&1
  echo "== 972 cluster $cl"; grep -E "walker_sweeps_per_s|k_inverse_cl" gpurun_out/r3a_q972_cl$cl.log | head -3
  python - gpurun_out/r3a_q972_cl$cl.log <<'PY'
import sys,re,json
t=open(sys.argv[1]).read()
i=t.find('{'); j=t.rfind('}')
try:
    d=json.loads(t[i:j+1]); print({k:(round(v["ms"],2),v["launches"]) for k,v in d["timers"].items()}, d["E_site"], d["n_singular"])
except Exception as e: print("ERR",e, t[-800:])
PY
done
for cl in 4 0; do
  timeout 300 python tools/quick_bench.py --n 12 --B 0.02 --walkers 4096 --sweeps 216 --therm 216 --opt inverse_cluster=$cl > gpurun_out/r3a_qc128_cl$cl.log 2>&1
  echo "== c128 432 cluster $cl"; grep -E "walker_sweeps_per_s" gpurun_out/r3a_qc128_cl$cl.log | head -3
  python - gpurun_out/r3a_qc128_cl$cl.log <<'PY'
import sys,re,json
t=open(sys.argv[1]).read()
i=t.find('{'); j=t.rfind('}')
try:
    d=json.loads(t[i:j+1]); print({k:(round(v["ms"],2),v["launches"]) for k,v in d["timers"].items()}, d["E_site"], d["n_singular"])
except Exception as e: print("ERR",e, t[-800:])
PY
done
