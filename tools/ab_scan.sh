#!/bin/bash
# GPU: time the scan-solver variants in lib/variants back to back on the default bench batch
# (variants named on the command line, plus those listed in tools/ab_extra.txt if present)
cd "$(dirname "$0")/.."
EXTRA=""
[ -f tools/ab_extra.txt ] && EXTRA=$(cat tools/ab_extra.txt)
for v in "$@" $EXTRA; do
  echo "=== $v"
  IBS_LIB=$PWD/ideal-ballooning-solver_b200/lib/variants/libibs_$v.so python tools/time_stages.py d3d 37 2>&1 | grep -E "solve|mean"
done
# piggy-backed on the same GPU call (boxes are scarce): geometry variants and the GPU test-suite
if [ -f tools/ab_piggyback.sh ]; then bash tools/ab_piggyback.sh; fi
