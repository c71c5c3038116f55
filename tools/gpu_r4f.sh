#!/bin/bash
mkdir -p gpurun_out
python bench.py --lattice 6 --steps 40 --warmup 3 --no-cpu-baseline --no-carlo --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['metric'], round(d['value']/1e6,2), d['ms_per_step'])"
OUT=gpurun_out/r4f_sanitizer_strict.txt
{
echo "\$ KDSL_LIB=libkdsl_strict.so (make STRICT=1: k_resident's re-evaluation loads predicated) compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k \"refresh_matches_oracle or imbalanced or complex_refresh or trajectory_bit_exact\""
KDSL_LIB=$PWD/kagomedsl.jl_b200/csrc/libkdsl_strict.so timeout 2400 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "refresh_matches_oracle or imbalanced or complex_refresh or trajectory_bit_exact" 2>&1 | tail -3
} > $OUT 2>&1
cat $OUT
