#!/bin/bash
# throughput-regime A/B of tuning builds: 4096 chains x 20000 epochs per kind (prof_kinds.py), and 4096 x 100 astrometry
for L in "$@"; do
  export OCTO_B200_LIB=octofitter.jl_b200/lib/libocto_$L.so
  echo "== $L"
  python profiles/tools/prof_kinds.py 4096 20000 5 2>&1 | tail -2 | sed 's/ms=\[[^]]*\]//'
  python profiles/tools/prof_kinds.py 4096 100 9 2>&1 | tail -2
done
