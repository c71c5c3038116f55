#!/bin/bash
mkdir -p gpurun_out
for fv in 0 4; do
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_flush -s 40 -c 40 --csv --log-file gpurun_out/r2o_flush_launches_fv$fv.csv python tools/quick_bench.py --n 12 --walkers 4096 --sweeps 432 --therm 432 --no-prof --opt flush_variant=$fv > gpurun_out/r2o_l$fv.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:k_flush_tma -s 20 -c 1 -o gpurun_out/prof_flushtma_r2o -f python tools/quick_bench.py --n 12 --walkers 4096 --sweeps 432 --therm 432 --no-prof --opt flush_variant=4 > gpurun_out/ncu_flushtma_r2o.log 2>&1
