#!/bin/bash
# what geometry() picks (auto) against a few forced candidates, on the calibration workloads and on long loops
[ -n "$1" ] && export OCTO_B200_LIB=octofitter.jl_b200/lib/libocto_$1.so
for W in C1 C2 C4 4096x10 4096x32 4096x100 1024x0+100 1024x10 2048x50; do
  python profiles/tools/sweep_geom.py $W 100 None 1,1,1 4,1,1 1,0,1 2,0,1 8,0,1 2>&1 | sed 's/ us\/step//'
done
for W in 4096x3162 4096x10000 4096x0+3162 4096x0+10000 4096x1500+1500; do
  python profiles/tools/sweep_geom.py $W 20 None 1,1,1 2>&1 | sed 's/ us\/step//'
  OCTO_B200_LAT_CAP=1000000 python profiles/tools/sweep_geom.py $W 20 None 2>&1 | sed 's/ us\/step//; s/force=None/cap=inf/'
done
