#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py -m gpu -x -q -k "complex or c128" > gpurun_out/r3n_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3n_pytest.log
tail -4 gpurun_out/r3n_pytest.log
timeout 300 python tools/quick_bench.py --n 12 --B 0.02 --walkers 4096 --sweeps 216 --therm 216 > gpurun_out/r3n_qc128.log 2>&1
python - gpurun_out/r3n_qc128.log <<'PY'
import sys,re,json
t=open(sys.argv[1]).read()
i=t.find('{'); j=t.rfind('}')
d=json.loads(t[i:j+1]); print(round(d["walker_sweeps_per_s"]/1e6,2), {k:(round(v["ms"],2),v["launches"]) for k,v in d["timers"].items()}, d["E_site"], d["n_singular"])
PY
bash tools/capture_profiles_r3.sh
