#!/bin/bash
# Round-2 ncu captures (runs ON THE GPU BOX under gpurun).  usage: capture_profiles_r2.sh A|B
# A: launch list of the bench command + --set full of k_reeval_fused, k_flush_wb, k_decide_wb (432 sites), k_resident (108)
# B: 972 sites (k_inverse_v4, k_gemm_W_dmma, k_flush_wb) + the ComplexF64 flush (432 sites, B != 0)
set -u
OUT=gpurun_out
mkdir -p $OUT
FULL="--set full --clock-control none --import-source on"
if [[ "${1:-A}" == "A" ]]; then
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 1200 --csv --log-file $OUT/launches_r2.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --thermalization 216 > $OUT/launches_r2.log 2>&1
ncu $FULL -k regex:k_reeval_fused -s 2 -c 1 -o $OUT/prof_fused_r2 -f python tools/prof_refresh.py 12 1024 3 > $OUT/ncu_fused_r2.log 2>&1
for spec in "k_flush_wb:flush:40" "k_decide_wb:decide:40"; do
    IFS=: read k n skip <<< "$spec"
    ncu $FULL -k regex:$k -s $skip -c 1 -o $OUT/prof_${n}_r2 -f python tools/quick_bench.py --walkers 4096 --sweeps 432 --therm 216 --no-prof > $OUT/ncu_${n}_r2.log 2>&1
done
ncu $FULL -k regex:k_resident -s 1 -c 1 -o $OUT/prof_resident_r2 -f python tools/quick_bench.py --n 6 --walkers 4096 --sweeps 216 --therm 54 --no-prof > $OUT/ncu_resident_r2.log 2>&1
else
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 600 --csv --log-file $OUT/launches972_r2.csv \
    python bench.py --lattice 18 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --thermalization 486 --walkers-per-gpu 2048 > $OUT/launches972_r2.log 2>&1
ncu $FULL -k regex:k_inverse_v4 -s 2 -c 1 -o $OUT/prof_inverse972_r2 -f python tools/prof_refresh.py 18 512 2 > $OUT/ncu_inverse972_r2.log 2>&1
ncu $FULL -k regex:k_gemm_W_dmma -s 1 -c 1 -o $OUT/prof_gemm972_r2 -f python tools/prof_refresh.py 18 512 2 > $OUT/ncu_gemm972_r2.log 2>&1
ncu $FULL -k regex:k_flush_wb -s 20 -c 1 -o $OUT/prof_flush972_r2 -f python tools/quick_bench.py --n 18 --walkers 2048 --sweeps 486 --therm 486 --no-prof > $OUT/ncu_flush972_r2.log 2>&1
ncu $FULL -k regex:k_flush_c -s 20 -c 1 -o $OUT/prof_flushc_r2 -f python tools/quick_bench.py --n 12 --walkers 4096 --sweeps 432 --therm 216 --B 0.02 --no-prof > $OUT/ncu_flushc_r2.log 2>&1
fi
ls -la $OUT | grep r2 | tail -20
