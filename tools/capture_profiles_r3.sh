#!/bin/bash
# Round-2 session-3 ncu captures (runs ON THE GPU BOX under gpurun): the ComplexF64 tensor-pipe kernels and the cluster inverse
set -u
OUT=gpurun_out
mkdir -p $OUT
FULL="--set full --clock-control none --import-source on"
QC="python tools/quick_bench.py --n 12 --walkers 4096 --sweeps 216 --therm 216 --B 0.02 --no-prof"
# usage: capture_profiles_r3.sh A|B   (two calls: gpurun copies at most 64 MiB back per call)
if [[ "${1:-A}" == "A" ]]; then
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 1200 --csv --log-file $OUT/launches_r3.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --thermalization 216 > $OUT/launches_r3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 600 --csv --log-file $OUT/launchesc128_r3.csv \
    python bench.py --B 0.02 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --thermalization 216 > $OUT/launchesc128_r3.log 2>&1
ncu $FULL -k regex:k_inverse_cl_c -s 0 -c 1 -o $OUT/prof_invclc_r3 -f $QC > $OUT/ncu_invclc_r3.log 2>&1
ncu $FULL -k regex:k_gemm_W_dmma_c -s 0 -c 1 -o $OUT/prof_gemmc_r3 -f $QC > $OUT/ncu_gemmc_r3.log 2>&1
else
ncu $FULL -k regex:k_flush_dmma_c -s 20 -c 1 -o $OUT/prof_flushdc_r3 -f $QC > $OUT/ncu_flushdc_r3.log 2>&1
ncu $FULL -k regex:k_inverse_cl -s 2 -c 1 -o $OUT/prof_invcl972_r3 -f python tools/prof_refresh.py 18 512 2 > $OUT/ncu_invcl972_r3.log 2>&1
fi
ls -la $OUT | grep r3 | tail -20
