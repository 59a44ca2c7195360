#!/bin/bash
# GPU: time the geometry-kernel variants in lib/variants on the NCSX and HBERG configs
cd "$(dirname "$0")/.."
for v in "$@"; do
  echo "=== $v"
  for wl in ncsx hberg; do
    IBS_LIB=$PWD/ideal-ballooning-solver_b200/lib/variants/libibs_$v.so python tools/time_stages.py $wl 1 2>&1 | grep -E "geometry"
  done
done
