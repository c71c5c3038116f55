#!/bin/bash
mkdir -p gpurun_out
export KDSL_LIB=$PWD/kagomedsl.jl_b200/csrc/libkdsl_ticks.so
for cl in 4 5 3; do
  echo "== 972 cluster $cl"; timeout 300 python tools/cl_phases.py 18 1024 inverse_cluster=$cl 2>&1 | tail -4
done
