#!/bin/bash
# 2-GPU: C-ABI NCCL group test + torchrun bench with the in-library all-reduce
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "two_gpus" > gpurun_out/r3x_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3x_pytest.log
tail -5 gpurun_out/r3x_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-carlo > gpurun_out/r3x_bench_2gpu.json 2> gpurun_out/r3x_bench_2gpu.err
tail -c 800 gpurun_out/r3x_bench_2gpu.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r3x_bench_2gpu.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("metric","value","n_gpus","ms_per_step")}, d.get("reduce"), (d.get("e2e") or {}).get("value"), d.get("observables"))
except Exception as e:
    print("ERR", e)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r3x_bench_ref_2gpu.json 2> gpurun_out/r3x_bench_ref_2gpu.err
tail -c 300 gpurun_out/r3x_bench_ref_2gpu.json
