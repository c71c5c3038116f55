#!/bin/bash
# compute-sanitizer over the session's kernels (cluster inverse / re-evaluation, ComplexF64 tensor-pipe kernels)
mkdir -p gpurun_out
OUT=gpurun_out/r3_sanitizer.txt
SEL="cluster or complex or c128"
{
echo "# compute-sanitizer on the round-2 final build; selection: pytest -k \"$SEL\" (k_inverse_cl, k_reeval_cl, k_inverse_cl_c, k_gemm_W_dmma_c, k_flush_dmma_c, k_unsplit_c, ...)"
echo "\$ compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py -m gpu -x -q -k \"$SEL\""
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py -m gpu -x -q -k "$SEL" 2>&1 | tail -4
echo "\$ compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k \"complex_cluster or cluster_reeval or complex_replay\""
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "complex_cluster or cluster_reeval or complex_replay" 2>&1 | tail -4
echo "\$ compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k \"complex_cluster or cluster_reeval or complex_replay\""
timeout 1500 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "complex_cluster or cluster_reeval or complex_replay" 2>&1 | tail -4
} > $OUT 2>&1
cat $OUT
