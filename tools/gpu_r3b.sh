#!/bin/bash
mkdir -p gpurun_out
export KDSL_LIB=$PWD/kagomedsl.jl_b200/csrc/libkdsl_ticks.so
for cl in 4 3 5; do
 for rs in 8 4; do
  echo "== 972 cluster $cl rs $rs"; timeout 300 python tools/cl_phases.py 18 1024 inverse_cluster=$cl inverse_row_slices=$rs 2>&1 | tail -4
 done
done
echo "== c128-size (n=12 via 2N embedding not available in real mode; use n=16: Np=384)"; timeout 300 python tools/cl_phases.py 16 1024 inverse_cluster=4 2>&1 | tail -4
