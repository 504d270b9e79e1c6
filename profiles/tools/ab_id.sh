#!/bin/bash
for L in "$@"; do
  export OCTO_B200_LIB=octofitter.jl_b200/lib/libocto_$L.so
  echo "== $L"; python profiles/tools/resident_diff.py 2>&1 | grep -v "n_diff 0"
done
