#!/bin/bash
# A/B of tuning builds (octofitter.jl_b200/lib/libocto_<name>.so, see build.py build_variant) in ONE box: per-leapfrog time
# of the resident explorer and back-to-back C2 steps.   bash profiles/tools/ab.sh base variant1 variant2 ... base
for L in "$@"; do
  export OCTO_B200_LIB=octofitter.jl_b200/lib/libocto_$L.so
  python profiles/tools/hmc_time.py 1024 50 10 2>&1 | grep "^lib"
  python profiles/tools/c2_steps.py C2
done
