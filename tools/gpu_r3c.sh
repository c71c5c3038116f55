#!/bin/bash
# round-2 session 3: cluster re-evaluation k_reeval_cl bring-up (parity first, then 972-site timing)
mkdir -p gpurun_out
export KDSL_DEBUG_OCC=1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cluster" > gpurun_out/r3c_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3c_pytest.log
tail -25 gpurun_out/r3c_pytest.log
run() {
  tag=$1; shift
  timeout 300 python tools/quick_bench.py --n 18 --walkers 2048 --sweeps 486 --therm 486 "$@" > gpurun_out/r3c_q972_$tag.log 2>&1
  echo "== 972 $tag"; grep -E "walker_sweeps_per_s|k_reeval_cl" gpurun_out/r3c_q972_$tag.log | head -3
  python - gpurun_out/r3c_q972_$tag.log <<'PY'
import sys,re,json
t=open(sys.argv[1]).read()
i=t.find('{'); j=t.rfind('}')
try:
    d=json.loads(t[i:j+1]); print({k:(round(v["ms"],2),v["launches"]) for k,v in d["timers"].items()}, d["E_site"], d["n_singular"])
except Exception as e: print("ERR",e, t[-800:])
PY
}
run v7
run v8_cl4_rs4 --opt inverse_variant=8
run v8_cl4_rs2 --opt inverse_variant=8 --opt reeval_rs=2
run v8_cl4_rs8 --opt inverse_variant=8 --opt reeval_rs=8
run v8_cl3_rs4 --opt inverse_variant=8 --opt reeval_cluster=3
run v8_cl5_rs4 --opt inverse_variant=8 --opt reeval_cluster=5
run v8_cl6_rs4 --opt inverse_variant=8 --opt reeval_cluster=6
