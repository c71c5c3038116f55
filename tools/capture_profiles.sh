#!/bin/bash
# Runs ON THE GPU BOX (under gpurun): the ncu launch list of the bench command and one `--set full` capture of each
# hot kernel.  Outputs go to gpurun_out/ (merged back by gpurun); tools/summarise_profiles.py turns them into
# the tracked summaries under profiles/.  Numbers printed under ncu are never bench values.
set -u
# usage: capture_profiles.sh TAG [names...]   (names from: launches fused inverse gemm gather flush decide measure; gpurun
# copies at most 64 MiB back per call, a full report is ~11 MiB)
TAG=${1:-r1}
shift
WANT=" ${*:-launches fused flush decide} "
OUT=gpurun_out
mkdir -p $OUT
if [[ "$WANT" == *" launches "* ]]; then
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 1200 --csv --log-file $OUT/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --thermalization 216 > $OUT/launches_${TAG}.log 2>&1
fi
if [[ "$WANT" == *" fused "* ]]; then       # the one-kernel re-evaluation (default): 1024 walkers = 2048 matrices per launch
    ncu --set full --clock-control none --import-source on -k regex:k_reeval_fused -s 2 -c 1 -o $OUT/prof_fused_${TAG} -f \
        python tools/prof_refresh.py 12 1024 3 > $OUT/ncu_fused_${TAG}.log 2>&1
fi
for spec in "k_inverse_v:inverse" "k_gemm_W_dmma:gemm" "k_gather_tilde:gather"; do
    k=${spec%%:*}; n=${spec##*:}
    [[ "$WANT" == *" $n "* ]] || continue
    ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o $OUT/prof_${n}_${TAG} -f \
        python tools/prof_refresh.py 12 1024 3 inverse_variant=5 > $OUT/ncu_${n}_${TAG}.log 2>&1
done
for spec in "k_flush_wb:flush:40" "k_decide_wb:decide:40" "k_measure_wb:measure:1"; do
    IFS=: read k n skip <<< "$spec"
    [[ "$WANT" == *" $n "* ]] || continue
    ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o $OUT/prof_${n}_${TAG} -f \
        python tools/quick_bench.py --walkers 4096 --sweeps 432 --therm 216 --no-prof > $OUT/ncu_${n}_${TAG}.log 2>&1
done
ls -la $OUT | tail -20
