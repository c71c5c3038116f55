#!/bin/bash
# compute-sanitizer over the FINAL build: the whole GPU suite under memcheck; racecheck / synccheck over the kernels touched last
mkdir -p gpurun_out
OUT=gpurun_out/r4_sanitizer.txt
{
echo "# compute-sanitizer on the final build of round 2 (k_reeval_fused<...,256> two CTAs per SM, wb_entry_warp2, k_gemm_W_dmma_c with unit rows, stream-ordered host copies)"
echo "\$ compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_headline.py::test_energy_432_sites_within_error_bars"
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_headline.py::test_energy_432_sites_within_error_bars 2>&1 | tail -4
echo "\$ compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k \"refresh_matches_oracle or imbalanced or complex_refresh or trajectory_bit_exact\""
timeout 2400 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "refresh_matches_oracle or imbalanced or complex_refresh or trajectory_bit_exact" 2>&1 | tail -4
echo "\$ compute-sanitizer --tool synccheck (same selection)"
timeout 2400 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "refresh_matches_oracle or imbalanced or complex_refresh or trajectory_bit_exact" 2>&1 | tail -4
} > $OUT 2>&1
cat $OUT
