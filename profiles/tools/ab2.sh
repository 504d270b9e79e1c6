#!/bin/bash
# like ab.sh plus the 4096 x 100 and 1024 x 1000 astrometry shapes
for L in "$@"; do
  export OCTO_B200_LIB=octofitter.jl_b200/lib/libocto_$L.so
  python profiles/tools/hmc_time.py 1024 50 10 2>&1 | grep "^lib"
  python profiles/tools/c2_steps.py C2
  python profiles/tools/sweep_geom.py 4096x100 100 None 2>&1 | tail -1
  python profiles/tools/sweep_geom.py 1024x1000 100 None 2>&1 | tail -1
done
