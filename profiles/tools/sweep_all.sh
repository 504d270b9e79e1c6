#!/bin/bash
# geometry sweeps of the workloads the cost model of geometry() was calibrated on, for tuning builds (see ab.sh)
for L in "$@"; do
  export OCTO_B200_LIB=octofitter.jl_b200/lib/libocto_$L.so
  for W in C2 4096x100 4096x1000 1024x1000 256x5000 C4 C3; do
    echo "== $L $W"
    python profiles/tools/sweep_geom.py $W 100 None 1,1,1 2,1,1 4,1,1 8,1,1 16,1,1 32,1,1 1,1,2 1,1,4 2,1,2 4,1,2 1,0,1 2,0,1 4,0,1 8,0,1 1,0,2 1,0,4 2,0,2 4,0,2 2>&1 | sed 's/ us\/step//'
  done
done
