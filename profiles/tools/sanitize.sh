#!/bin/bash
# compute-sanitizer over profiles/tools/sanitize_run.py (every kernel mode): memcheck, racecheck, initcheck
for T in memcheck racecheck initcheck; do
  echo "=== $T"
  compute-sanitizer --tool $T python profiles/tools/sanitize_run.py 2>&1 | grep -v "^$" | tail -60
done
