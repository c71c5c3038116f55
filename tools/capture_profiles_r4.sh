#!/bin/bash
# last ncu captures of round 2: the ComplexF64 product (now writing the unit rows) and flush (four tiles in flight), k_reeval_fused<..., 256> at 192 sites
set -u
OUT=gpurun_out
mkdir -p $OUT
FULL="--set full --clock-control none --import-source on"
QC="python tools/quick_bench.py --n 12 --walkers 4096 --sweeps 216 --therm 216 --B 0.02 --no-prof"
ncu $FULL -k regex:k_gemm_W_dmma_c -s 0 -c 1 -o $OUT/prof_gemmc_r4 -f $QC > $OUT/ncu_gemmc_r4.log 2>&1
ncu $FULL -k regex:k_flush_dmma_c -s 20 -c 1 -o $OUT/prof_flushdc_r4 -f $QC > $OUT/ncu_flushdc_r4.log 2>&1
ncu $FULL -k regex:k_reeval_fused -s 2 -c 1 -o $OUT/prof_fused192_r4 -f python tools/prof_refresh.py 8 2048 3 > $OUT/ncu_fused192_r4.log 2>&1
ls -la $OUT | grep r4 | tail
