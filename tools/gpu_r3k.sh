#!/bin/bash
mkdir -p gpurun_out
export KDSL_DEBUG_OCC=1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cluster" > gpurun_out/r3k_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3k_pytest.log
tail -8 gpurun_out/r3k_pytest.log
run() {
  tag=$1; shift
  timeout 300 python tools/quick_bench.py --n 12 --walkers 4096 --sweeps 432 --therm 432 "$@" > gpurun_out/r3k_q432_$tag.log 2>&1
  echo "== 432 $tag"; grep -E "k_reeval_cl" gpurun_out/r3k_q432_$tag.log | head -1
  python - gpurun_out/r3k_q432_$tag.log <<'PY'
import sys,re,json
t=open(sys.argv[1]).read()
i=t.find('{'); j=t.rfind('}')
try:
    d=json.loads(t[i:j+1]); print(round(d["walker_sweeps_per_s"]/1e6,2), {k:(round(v["ms"],2),v["launches"]) for k,v in d["timers"].items()}, d["E_site"], d["n_singular"], round(d["flush_GBs"]))
except Exception as e: print("ERR",e, t[-800:])
PY
}
run fused
run cl4 --opt inverse_variant=8
run cl4_t512 --opt inverse_variant=8 --opt inverse_tuning=1
run cl3 --opt inverse_variant=8 --opt reeval_cluster=3
run cl2 --opt inverse_variant=8 --opt reeval_cluster=2
run cl4_rs2 --opt inverse_variant=8 --opt reeval_rs=2
run cl4_rs1 --opt inverse_variant=8 --opt reeval_rs=1
run cl3_rs2 --opt inverse_variant=8 --opt reeval_cluster=3 --opt reeval_rs=2
run cl2_rs2 --opt inverse_variant=8 --opt reeval_cluster=2 --opt reeval_rs=2
