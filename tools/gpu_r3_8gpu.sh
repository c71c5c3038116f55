#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 --no-carlo > gpurun_out/r3_bench_8gpu.json 2> gpurun_out/r3_bench_8gpu.err
tail -c 400 gpurun_out/r3_bench_8gpu.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r3_bench_8gpu.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("metric","value","n_gpus","ms_per_step")}, d.get("reduce"), (d.get("e2e") or {}).get("value"), d.get("observables"))
except Exception as e:
    print("ERR", e)
PY
